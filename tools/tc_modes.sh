#!/bin/bash
# Correctness (vs bf16 emulation / fp32 reference) and speed of the tensor-core tower variants.
for mode in ${MODES:-0 2}; do
  echo "=== AZ_TC_MODE=$mode"
  AZ_TC_MODE=$mode timeout 300 python tools/dbg_net.py bf16 2>&1 | tail -4
  AZ_TC_MODE=$mode timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --games ${GAMES:-4096} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sims/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'tower', d['roofline']['note'].split('last tick (')[1][:10], 'frac', round(d['roofline']['frac'],3))"
done
python tools/dbg_net.py fp32 2>&1 | tail -4
