#!/bin/bash
echo "=== AZ_TC_MODE=4 (2-CTA pair)"
AZ_TC_MODE=4 timeout 120 python tools/dbg_net.py bf16 2>&1 | tail -4
echo "rc=$?"
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
AZ_TC_MODE=4 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sims/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'tower', d['roofline']['note'].split('last tick (')[1][:10], 'frac', round(d['roofline']['frac'],3))"
AZ_TC_MODE=2 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('mode2 sims/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'tower', d['roofline']['note'].split('last tick (')[1][:10], 'frac', round(d['roofline']['frac'],3))"
