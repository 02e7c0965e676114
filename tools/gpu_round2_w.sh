#!/bin/bash
# weight multicast in clusters of two (single-CTA resident build): tests, conv timings on / off, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -p no:cacheprovider -k "statistics or concat or qconv or conv" > gpurun_out/w_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/w_tests.log
for m in 0 1; do
echo "--- mcast $m"; ATTNDM_TC_MCAST=$m timeout 300 python tools/conv_bench.py --shapes c128_32,c128_16,c128_8,out_32,in_32 --stats 0 2>&1
echo "--- mcast $m +stats"; ATTNDM_TC_MCAST=$m timeout 300 python tools/conv_bench.py --shapes c128_32,c128_16,c128_8 --stats 1 2>&1
done
ATTNDM_TC_MCAST=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/w_bench.json').read().strip().splitlines()[-1])
print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', d.get('launches_per_denoising_step'))
PY
