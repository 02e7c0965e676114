#!/bin/bash
# ncu --set full of the dominant conv kernel and of the HBM-bound kernels (each after the same command ran clean)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --graph 0 --shapes c128_32 --iters 3 > gpurun_out/conv_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv_r02 python tools/conv_bench.py --graph 0 --shapes c128_32 --iters 3 > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv rc=$?"
timeout 300 python tools/hbm_kernels.py 5 > gpurun_out/hbm_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"act_quant_rows|gn_stats_kernel|minmax_partial|calib_mix|ddim_step" -c 10 \
    -f -o gpurun_out/prof_hbm_r02 python tools/hbm_kernels.py 1 > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
cat gpurun_out/hbm_plain.log
