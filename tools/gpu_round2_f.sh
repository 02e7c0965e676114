#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --shapes c128_32 --iters 3 > gpurun_out/conv_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_r02a python tools/conv_bench.py --shapes c128_32 --iters 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu conv rc=$?"
