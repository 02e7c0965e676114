#!/bin/bash
# ncu launch list (device time of every kernel) for one eager UNet forward, CIFAR config, batch 256.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py --steps 1 --events 0 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 1500 --csv \
   --log-file gpurun_out/launches_r01.csv python tools/profile_step.py --steps 1 --events 0 > gpurun_out/ncu_run.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r01.csv
