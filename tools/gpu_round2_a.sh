#!/bin/bash
# GPU call A (round 2): changed unit tests, lock-step parity at full size for both SiLU builds, step time A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/parity
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/a_env.log 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/a_env.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/a_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/a_tests.log
ATTNDM_PARITY_OUT=gpurun_out/parity timeout 1500 python -m pytest tests/test_gpu_lockstep.py -q -s -m gpu -p no:cacheprovider > gpurun_out/a_lockstep.log 2>&1
echo "lockstep rc=$?" | tee -a gpurun_out/a_lockstep.log
ATTNDM_LIB=$PWD/attentiondm_b200/libattndm_b200_silu_sfu.so ATTNDM_PARITY_OUT=gpurun_out/parity timeout 900 python -m pytest tests/test_gpu_lockstep.py -q -s -m gpu -p no:cacheprovider -k "tiny or cifar10" > gpurun_out/a_lockstep_sfu.log 2>&1
echo "lockstep sfu rc=$?" | tee -a gpurun_out/a_lockstep_sfu.log
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
ATTNDM_LIB=$PWD/attentiondm_b200/libattndm_b200_silu_sfu.so timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/a_bench_sfu.json 2> gpurun_out/a_bench_sfu.err
tail -3 gpurun_out/a_tests.log; grep -E "lockstep|passed|failed" gpurun_out/a_lockstep.log | tail -12; grep -E "lockstep|passed|failed" gpurun_out/a_lockstep_sfu.log | tail -6
python -c "
import json
for f in ('a_bench.json','a_bench_sfu.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'failed', e)
"
