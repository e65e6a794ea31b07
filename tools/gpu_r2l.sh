#!/bin/bash
# round 2, session l: coalesced heads kernel: network tests + bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_net_layers.py tests/test_gpu_engine.py tests/test_gpu_facade.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2l.err | tee gpurun_out/bench_r2l.json | cut -c1-200
timeout -s KILL 300 $B --workload gomoku13_c4 2>gpurun_out/bench_r2l_c4.err | tee gpurun_out/bench_r2l_c4.json | cut -c1-200
