#!/bin/bash
# ncu --set full of the halo conv kernel, plain and with epilogue statistics (source-level stall samples)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --graph 0 --shapes c128_32 --iters 3 --stats 1 > gpurun_out/t_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_t0 python tools/conv_bench.py --graph 0 --shapes c128_32 --iters 3 --stats 0 > gpurun_out/t_ncu0.log 2>&1
echo "ncu plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_t1 python tools/conv_bench.py --graph 0 --shapes c128_32 --iters 3 --stats 1 > gpurun_out/t_ncu1.log 2>&1
echo "ncu stats rc=$?"
ls -la gpurun_out/prof_conv_t*
