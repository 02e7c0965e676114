#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python bench.py --steps 3 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_last', d['e2e_keep_last']['value'])"
