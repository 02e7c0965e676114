#!/bin/bash
# L2-residency probe: the CIFAR step at batch 32 / 64 / 128 / 256 per GPU (images/s per batch size)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for b in 32 64 128 256; do
  timeout 600 python bench.py --steps 2 --warmup 3 --no-extras --batch $b > gpurun_out/o_bench_b$b.json 2> gpurun_out/o_bench_b$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/o_bench_b$b.json').read().strip().splitlines()[-1])
print($b, round(d['value'],1), 'img/s', round(d['ms_per_step']/100,3), 'ms/step', 'per-image us', round(d['ms_per_step']*10/$b,2))
PY
done
