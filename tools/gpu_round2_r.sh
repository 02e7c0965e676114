#!/bin/bash
# residual loads per whole block again (with / without the L2 prefetch of the residual rows), conv timings, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -p no:cacheprovider -k "statistics or concat or qconv or conv" > gpurun_out/r_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/r_tests.log
echo "--- residual prefetch off"; ATTNDM_TC_RES_PREFETCH=0 timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8 --stats 0 2>&1 | grep "res=1"
echo "--- residual prefetch off +stats"; ATTNDM_TC_RES_PREFETCH=0 timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8 --stats 1 2>&1 | grep "res=1"
echo "--- stats 0"; timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,c128_64 --stats 0 2>&1 | tee gpurun_out/r_conv0.log
echo "--- stats 1"; timeout 300 python tools/conv_bench.py --shapes c128_32,c256_32,c128_16,c128_8,c128_64 --stats 1 2>&1 | tee gpurun_out/r_conv1.log
for v in "1 1"; do
  set -- $v
  ATTNDM_CONV_STATS=$1 ATTNDM_TC_RES_PREFETCH=$2 timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r_bench_$1$2.json 2> gpurun_out/r_bench_$1$2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r_bench_$1$2.json').read().strip().splitlines()[-1])
print('conv_stats=$1 res_prefetch=$2', round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', d.get('launches_per_denoising_step'))
PY
done
