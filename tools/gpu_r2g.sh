#!/bin/bash
# round 2, session g: select kernel after the observation / prefetch changes; ncu --set full of the 64-filter tower kernel (C4)
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2g.err | tee gpurun_out/bench_r2g.json | cut -c1-200
timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2g_g512.err | tee gpurun_out/bench_r2g_g512.json | cut -c1-200
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_x -s 2900 -c 2 -o gpurun_out/prof_conv_c4_r2g -f python bench.py --workload gomoku13_c4 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_conv_c4_r2g.log 2>&1
tail -1 gpurun_out/ncu_conv_c4_r2g.log | cut -c1-200
timeout -s KILL 400 python -m pytest tests -m gpu -q -x -k "selfplay or traces or corpus or concurrent or stagger" 2>&1 | tail -3
