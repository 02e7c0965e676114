#!/bin/bash
# attention with R query rows per warp: tests, church / CelebA / CIFAR-W4 bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "attention or large_configs or mixed or ablation or calib" > gpurun_out/aa_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/aa_tests.log
for c in church_w8a8 celeba_w8a8 cifar10_w4_attn cifar10_w8a8; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 --no-extras > gpurun_out/aa_bench_$c.json 2> gpurun_out/aa_bench_$c.err
  python - <<PY
import json
e=json.loads(open('gpurun_out/aa_bench_$c.json').read().strip().splitlines()[-1])
print('$c', round(e['value'],1), 'img/s', round(e['ms_per_step'],1), 'ms/pass', e['launches_per_denoising_step'], 'launches', 'e2e', round(e['e2e']['value'],1))
PY
done
