#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "conv" 2>&1 | tail -3 > gpurun_out/e_tests.log
cat gpurun_out/e_tests.log
for dbg in 0 2 64 66 96; do echo "DBG=$dbg $(ATTNDM_TC_DBG=$dbg timeout 120 python tools/conv_bench.py --shapes c128_32 2>&1 | grep 'res=0')"; done > gpurun_out/e_var.log 2>&1
echo "split=0 $(ATTNDM_TC_SPLIT=0 timeout 120 python tools/conv_bench.py --shapes c128_32 2>&1 | grep 'res=0')" >> gpurun_out/e_var.log
cat gpurun_out/e_var.log
timeout 200 python tools/conv_bench.py 2>&1 | tee gpurun_out/e_conv_bench.log
export ATTNDM_LIB=$PWD/attentiondm_b200/libattndm_b200_tc_trace.so
for dbg in 0 66; do
  echo "=== ATTNDM_TC_DBG=$dbg"
  ATTNDM_TC_DBG=$dbg TRACE_CTA=5 TRACE_ITS=14 timeout 120 python tools/conv_trace.py c128_32 2>&1 | tail -130
done > gpurun_out/e_trace.log 2>&1
