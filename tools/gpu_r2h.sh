#!/bin/bash
# round 2, session h: early accumulator release + vector bias loads in the conv epilogue: per-layer parity, C2 / C4 / 512-game benches
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_net_layers.py -m gpu -q -x 2>&1 | tail -4
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2h.err | tee gpurun_out/bench_r2h.json | cut -c1-200
timeout -s KILL 300 $B --workload gomoku13_c4 2>gpurun_out/bench_r2h_c4.err | tee gpurun_out/bench_r2h_c4.json | cut -c1-200
timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2h_g512.err | tee gpurun_out/bench_r2h_g512.json | cut -c1-200
