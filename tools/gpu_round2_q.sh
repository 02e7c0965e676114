#!/bin/bash
# conv-epilogue GroupNorm statistics: tests, conv timings with / without, bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -p no:cacheprovider -k "statistics or concat or qconv" > gpurun_out/q_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/q_tests.log
echo "--- residual prefetch off"; ATTNDM_TC_RES_PREFETCH=0 timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8 --stats 0 2>&1 | grep "res=1"
echo "--- stats 0"; timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,c128_64 --stats 0 2>&1 | tee gpurun_out/q_conv0.log
echo "--- stats 1"; timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,c128_64 --stats 1 2>&1 | tee gpurun_out/q_conv1.log
for v in 1 0; do
  ATTNDM_CONV_STATS=$v timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/q_bench_$v.json 2> gpurun_out/q_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/q_bench_$v.json').read().strip().splitlines()[-1])
print('conv_stats=$v', round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', d.get('launches_per_denoising_step'))
PY
done
