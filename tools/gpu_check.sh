#!/bin/bash
# Standard GPU check: all GPU tests (with -s so the measured parity values are on disk), conv timings, the default bench line.
# usage: tools/gpu_check.sh <tag>
cd "$(dirname "$0")/.."
T=${1:-x}
mkdir -p gpurun_out
ATTNDM_PARITY_OUT=gpurun_out/parity_$T timeout 1500 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider > gpurun_out/${T}_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/${T}_tests.log
timeout 300 python tools/conv_bench.py > gpurun_out/${T}_conv_bench.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_conv_bench.log
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['ms_per_launch'])
for r in d.get('roofline_hbm', []): print(r['kernel'][:50], round(r['achieved']), round(r['frac'],3))
PY
