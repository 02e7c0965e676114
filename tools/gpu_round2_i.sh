#!/bin/bash
# bench lines for BASELINE.json configs 3, 4, 5 (one GPU) and the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in cifar10_w4_attn celeba_w8a8 church_w8a8; do
  timeout 1200 python bench.py --config $c --steps 2 --warmup 3 --no-extras > gpurun_out/i_bench_$c.json 2> gpurun_out/i_bench_$c.err
  echo "$c rc=$?"; tail -c 600 gpurun_out/i_bench_$c.err
done
python - <<'PY'
import json
for c in ('cifar10_w4_attn','celeba_w8a8','church_w8a8'):
    try:
        d=json.loads(open(f'gpurun_out/i_bench_{c}.json').read().strip().splitlines()[-1]); print(c, d['value'], d['unit'], d['ms_per_step'], d['config'].get('int8_layers'), d.get('e2e',{}).get('value'), d.get('whole_step_conv_tops'))
    except Exception as e: print(c, 'failed', e)
PY
