#!/bin/bash
# round 2, session b: per-layer parity of every conv kernel variant, trained-checkpoint parity, split-bf16 tower
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_gpu_net_layers.py -m gpu -q -s 2>&1 | tail -80 | tee gpurun_out/pytest_layers_r2b.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_engine.py -m gpu -q -x -k "net or search_with" 2>&1 | tail -15 | tee gpurun_out/pytest_net_r2b.txt
timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_r2b.txt
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision bf16x3 2>gpurun_out/bench_r2b_x3.err | tee gpurun_out/bench_r2b_x3.json | cut -c1-300
