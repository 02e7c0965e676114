#!/bin/bash
# in-place concat reads: GPU tests, bench with and without (ATTNDM_CAT=0)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/p_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/p_tests.log
for cat in 1 0; do
  ATTNDM_CAT=$cat timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/p_bench_cat$cat.json 2> gpurun_out/p_bench_cat$cat.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/p_bench_cat$cat.json').read().strip().splitlines()[-1])
print('cat=$cat', round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', d.get('launches_per_denoising_step'))
PY
done
