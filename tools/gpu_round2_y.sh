#!/bin/bash
# rowprog: hooks compiled out, B fragments preloaded, time embedding fetched early -- bit-identity tests and bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -p no:cacheprovider -k "rowprog or fused or cifar or engine or graph" > gpurun_out/y_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/y_tests.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/y_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', d.get('launches_per_denoising_step'))
PY
