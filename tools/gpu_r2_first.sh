#!/bin/bash
# First GPU session of round 2: what round 1 prepared but could not measure (its GPU minutes ran out).
#  1. full suite, single process (what the driver runs) and the smoke
#  2. bench: default, AZ_PIPELINE=1 (expected: tick ~= network phase), cold start for continuity with the round-1 numbers
#  3. other configs on the staggered population (C4 gomoku, C5 19x19), 2-GPU scaling if a 2-GPU box is available
#  4. ncu: launch list of one full step on the staggered population is too slow under ncu (7000 prologue launches): use --cold-start
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2a.txt
timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout -s KILL 240 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_r2a.err | tee gpurun_out/bench_r2a.json | cut -c1-300
AZ_PIPELINE=1 timeout -s KILL 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2a_pipeline.err | tee gpurun_out/bench_r2a_pipeline.json | cut -c1-300
timeout -s KILL 240 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --cold-start 2>gpurun_out/bench_r2a_cold.err | tee gpurun_out/bench_r2a_cold.json | cut -c1-300
timeout -s KILL 240 python bench.py --workload gomoku13_c4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2a_c4.err | tee gpurun_out/bench_r2a_c4.json | cut -c1-300
timeout -s KILL 300 python bench.py --workload go19_c5 --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2a_c5.err | tee gpurun_out/bench_r2a_c5.json | cut -c1-300
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 140 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cold-start > gpurun_out/bench_under_ncu_r2a.log 2>&1
