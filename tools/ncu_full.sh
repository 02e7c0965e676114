#!/bin/bash
# one `ncu --set full` capture of the dominant kernel (128->128 3x3 @32x32, batch 256), after the same
# command ran clean without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/conv_bench.py --shapes c128_32 --iters 3 > gpurun_out/conv_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qconv_i8 -s 3 -c 1 \
    -f -o gpurun_out/prof_conv python tools/conv_bench.py --shapes c128_32 --iters 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
