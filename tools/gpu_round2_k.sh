#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "conv" -p no:cacheprovider 2>&1 | tail -4
for pr in 0 1; do echo "PAIR=$pr"; ATTNDM_TC_PAIR=$pr timeout 200 python tools/conv_bench.py --shapes c128_32,c128_64,c128_128,c128_256 2>&1; done | tee gpurun_out/k_pair.log
ATTNDM_PARITY_OUT=gpurun_out/parity_k timeout 900 python -m pytest tests/test_gpu_lockstep.py -q -s -p no:cacheprovider -k "celeba or church" 2>&1 | grep -E "^\[lockstep|passed|failed" | cut -c1-400
