#!/bin/bash
cd "$(dirname "$0")/.."
c=${1:-church_w8a8}
export ATTNDM_CONFIG=$c
timeout 600 python tools/profile_engine.py --steps 1 --events 0 > gpurun_out/ncu_plain_$c.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off -c 2500 --csv \
   --log-file gpurun_out/launches_r02_${c}_final.csv python tools/profile_engine.py --steps 1 --events 0 > gpurun_out/ncu_run_$c.log 2>&1
echo "ncu $c rc=$?"
python tools/parse_launches.py gpurun_out/launches_r02_${c}_final.csv > gpurun_out/launches_r02_${c}_final.txt
head -60 gpurun_out/launches_r02_${c}_final.txt
