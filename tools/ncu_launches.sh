#!/bin/bash
# ncu launch list (device time of every kernel) of ONE engine step (the body the CUDA graph captures: staged
# tables, fused rowprog launches, layer kernels, DDIM update), CIFAR config, batch 256.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r01}
timeout 600 python tools/profile_engine.py --steps 1 --events 0 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 1500 --csv \
   --log-file gpurun_out/launches_$R.csv python tools/profile_engine.py --steps 1 --events 0 > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_$R.csv
python tools/parse_launches.py gpurun_out/launches_$R.csv > gpurun_out/launches_$R.txt
head -45 gpurun_out/launches_$R.txt
