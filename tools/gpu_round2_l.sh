#!/bin/bash
cd "$(dirname "$0")/.."
for c in celeba_w8a8 church_w8a8; do
  timeout 1500 python bench.py --config $c --steps 2 --warmup 3 --no-extras > gpurun_out/l_bench_$c.json 2> gpurun_out/l_bench_$c.err
  echo "$c rc=$?"; tail -c 300 gpurun_out/l_bench_$c.err
done
python - <<'PY'
import json
for c in ('celeba_w8a8','church_w8a8'):
    try:
        d=json.loads(open(f'gpurun_out/l_bench_{c}.json').read().strip().splitlines()[-1]); print(c, d['value'], d['unit'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('whole_step_conv_tops'))
    except Exception as e: print(c, 'failed', e)
PY
nvidia-smi --query-gpu=memory.used --format=csv
