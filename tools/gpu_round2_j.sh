#!/bin/bash
cd "$(dirname "$0")/.."
for rep in 1 2 3; do for dbg in 0 1024; do
  echo "DBG=$dbg $(ATTNDM_TC_DBG=$dbg timeout 120 python tools/conv_bench.py --shapes c128_32,c128_16,c128_8 2>&1 | grep 'res=0' | awk '{printf "%s %s us | ", $1, $3}')"
done; done | tee gpurun_out/j_tap.log
