#!/bin/bash
# Round-2 record: every GPU test with -s (measured parity values on disk), conv timings, the default bench line
# (rooflines, CUDA-eager and CPU baselines), the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ATTNDM_PARITY_OUT=gpurun_out/parity_r02 timeout 1800 python -m pytest tests -q -s -m gpu -p no:cacheprovider > gpurun_out/r02_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/r02_tests.log
timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,n256_32,out_32,in_32,c128_64,c128_128,c128_256 > gpurun_out/r02_conv_bench.log 2>&1
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "ref rc=$?"
tail -3 gpurun_out/r02_tests.log; cat gpurun_out/r02_conv_bench.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline'].get('ms_per_launch_eager'))
for r in d.get('roofline_hbm', []): print(r['kernel'][:50], round(r['achieved']), round(r['frac'],3))
print(d['cuda_eager_baseline']); print(d['cpu_baseline'])
print(open('gpurun_out/r02_bench_ref.json').read()[:600])
PY
