#!/bin/bash
# the whole GPU suite, conv timings with statistics, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/z_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/z_tests.log
timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,c128_64 --stats 1 2>&1 | tee gpurun_out/z_conv1.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/z_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', d.get('launches_per_denoising_step'))
PY
