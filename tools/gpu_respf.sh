#!/bin/bash
# A/B of the residual L2 prefetch (AZ_TC_RESPF) in the pair conv kernel: off, on, off, on; then the network / self-play parity
# tests with it on.
mkdir -p gpurun_out
for i in 1 2; do
  for v in 0 1; do
    AZ_TC_RESPF=$v timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>gpurun_out/respf.err | tee gpurun_out/bench_respf${v}_$i.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('RESPF=$v run $i', round(d['value']), 'sims/s  conv', round(d['roofline']['achieved'], 1), 'TF/s  clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
    tail -1 gpurun_out/respf.err
  done
done
AZ_TC_RESPF=1 timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "net or selfplay or search_with_cuda" 2>&1 | tail -3
