#!/bin/bash
# `ncu --set full` captures, each after the same command ran clean without ncu:
#   (1) the dominant kernel (qconv_i8_halo_kernel, 128->128 3x3 @32x32, batch 256, no residual)
#   (2) the fused trunk program (rowprog_kernel, one engine step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --shapes c128_32 --iters 3 > gpurun_out/conv_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8_halo -s 3 -c 1 \
    -f -o gpurun_out/prof_conv python tools/conv_bench.py --shapes c128_32 --iters 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu conv rc=$?"
timeout 300 python tools/profile_engine.py --events 0 --steps 1 > gpurun_out/engine_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rowprog_kernel -s 5 -c 1 \
    -f -o gpurun_out/prof_rowprog python tools/profile_engine.py --events 0 --steps 1 > gpurun_out/ncu_rowprog.log 2>&1
echo "ncu rowprog rc=$?"
