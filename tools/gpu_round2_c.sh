#!/bin/bash
# GPU call C: unit tests after the warp-role reorder / history rings / SFU default; conv timings; bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/c_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/c_tests.log
timeout 300 python tools/conv_bench.py > gpurun_out/c_conv_bench.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench rc=$?"
tail -3 gpurun_out/c_tests.log; cat gpurun_out/c_conv_bench.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_last', d['e2e_keep_last']['value'], 'roof', d['roofline']['frac'], d['roofline']['ms_per_launch'])
print(d['int8_peak']); print(d['cuda_eager_baseline']); print(d['cpu_baseline'])
for r in d['roofline_hbm']: print(r['kernel'][:50], round(r['achieved']), round(r['frac'],3))
PY
