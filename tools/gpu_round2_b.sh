#!/bin/bash
# GPU call B (round 2): guarded-SiLU tests, lock-step parity for the three SiLU builds, step-time A/B, the new
# bench line, reference arm, ncu --set full of the HBM-bound kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/parity
L=$PWD/attentiondm_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider > gpurun_out/b_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/b_tests.log
ATTNDM_PARITY_OUT=gpurun_out/parity timeout 1500 python -m pytest tests/test_gpu_lockstep.py -q -s -m gpu -p no:cacheprovider > gpurun_out/b_lockstep.log 2>&1
echo "lockstep rc=$?" | tee -a gpurun_out/b_lockstep.log
for v in silu_sfu silu_accurate; do
  ATTNDM_LIB=$L/libattndm_b200_$v.so ATTNDM_PARITY_OUT=gpurun_out/parity ATTNDM_LOCKSTEP_B=8 timeout 900 python -m pytest tests/test_gpu_lockstep.py -q -s -m gpu -p no:cacheprovider -k "cifar10" > gpurun_out/b_lockstep_$v.log 2>&1
  echo "lockstep $v rc=$?"
done
ATTNDM_PARITY_OUT=gpurun_out/parity_b8 ATTNDM_LOCKSTEP_B=8 timeout 900 python -m pytest tests/test_gpu_lockstep.py -q -s -m gpu -p no:cacheprovider -k "cifar10" > gpurun_out/b_lockstep_default_b8.log 2>&1
for v in default silu_sfu silu_accurate; do
  if [ $v = default ]; then unset ATTNDM_LIB; else export ATTNDM_LIB=$L/libattndm_b200_$v.so; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/b_bench_$v.json 2> gpurun_out/b_bench_$v.err
done
unset ATTNDM_LIB
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/b_bench_full.json 2> gpurun_out/b_bench_full.err
echo "bench full rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err
timeout 300 python tools/hbm_kernels.py 5 > gpurun_out/b_hbm_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"act_quant_rows|gn_stats_kernel|minmax_partial|calib_mix" -c 8 \
    -f -o gpurun_out/prof_hbm python tools/hbm_kernels.py 1 > gpurun_out/b_ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
tail -3 gpurun_out/b_tests.log; grep -hE "^\[lockstep|passed|failed" gpurun_out/b_lockstep*.log | cut -c1-600
python - <<'PY'
import json
for f in ('b_bench_default.json','b_bench_silu_sfu.json','b_bench_silu_accurate.json','b_bench_full.json','b_bench_ref.json'):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('e2e'))
    except Exception as e: print(f, 'failed', e)
PY
cat gpurun_out/b_hbm_plain.log
