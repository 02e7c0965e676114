#!/bin/bash
# what bounds the plain 128->128 conv now: debug switches (timing only) and the CTA-pair build
cd "$(dirname "$0")/.."
for d in 0 128 256 384 1 2 64 32 16; do
  echo "--- ATTNDM_TC_DBG=$d"; ATTNDM_TC_DBG=$d timeout 120 python tools/conv_bench.py --shapes c128_32,c128_16 --stats 0 2>&1 | grep "res=0"
done
echo "--- ATTNDM_TC_PAIR=2"; ATTNDM_TC_PAIR=2 timeout 120 python tools/conv_bench.py --shapes c128_32,c128_16,c128_8 --stats 0 2>&1
echo "--- ATTNDM_TC_PAIR=2 stats"; ATTNDM_TC_PAIR=2 timeout 120 python tools/conv_bench.py --shapes c128_32,c128_16,c128_8 --stats 1 2>&1
