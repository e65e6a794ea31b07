#!/bin/bash
# round 2, session j: input-layer K steps, batched heads loads: per-layer parity + benches
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_net_layers.py tests/test_gpu_engine.py -m gpu -q -x -k "layer or net or trained" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2j.err | tee gpurun_out/bench_r2j.json | cut -c1-200
timeout -s KILL 300 $B --workload gomoku13_c4 2>gpurun_out/bench_r2j_c4.err | tee gpurun_out/bench_r2j_c4.json | cut -c1-200
