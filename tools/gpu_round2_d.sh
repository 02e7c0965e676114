#!/bin/bash
cd "$(dirname "$0")/.."
export ATTNDM_LIB=$PWD/attentiondm_b200/libattndm_b200_tc_trace.so
for dbg in 0 2 1 32; do
  echo "=== ATTNDM_TC_DBG=$dbg" 
  ATTNDM_TC_DBG=$dbg TRACE_CTA=5 TRACE_ITS=8 timeout 120 python tools/conv_trace.py c128_32 2>&1 | tail -120
done > gpurun_out/d_trace.log 2>&1
unset ATTNDM_LIB
for dbg in 0 1 2 3 32 16 144; do echo "DBG=$dbg"; ATTNDM_TC_DBG=$dbg timeout 120 python tools/conv_bench.py --shapes c128_32,c128_16 2>&1 | grep "res=0"; done > gpurun_out/d_dbg.log 2>&1
cat gpurun_out/d_dbg.log
