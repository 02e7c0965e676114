#!/bin/bash
# Run every GPU test function in its own process (a CUDA fault in one kernel must not
# poison the context for the rest) and collect the logs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
FUNCS=$(grep -oE "^def (test_[a-z0-9_]+)" tests/test_gpu_parity.py | awk '{print $2}')
PASS=0; FAIL=0
for f in $FUNCS; do
  echo "=== $f" >> $LOG
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$f" -p no:cacheprovider --tb=short >> $LOG 2>&1
  rc=$?
  echo "=== $f rc=$rc" >> $LOG
  if [ $rc -eq 0 ]; then PASS=$((PASS+1)); else FAIL=$((FAIL+1)); echo "FAILED: $f (rc=$rc)"; fi
done
echo "functions passed=$PASS failed=$FAIL" | tee -a $LOG
grep -E "passed|failed|error" $LOG | tail -40
