#!/bin/bash
# Round-1 third session, ONE GPU call (budget ~15 min): validate the node-cache search + restart/profile ABI (full GPU suite), bench on the
# age-staggered population with the per-phase tick split, A/B against the replay search, then the experimental dense-x conv variant.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
date +%T
timeout -s KILL 480 python -m pytest tests -m gpu -q -n 4 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r1c.txt
date +%T
timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_r1c.txt
timeout -s KILL 240 python bench.py --steps 4 --warmup 3 2>gpurun_out/bench_r1c.err | tee gpurun_out/bench_r1c.json | cut -c1-600
tail -2 gpurun_out/bench_r1c.err
date +%T
for tag in go9_small64 go9_c2 gomoku13_c4; do
  AZ_TC_MODE=5 timeout -s KILL 90 python tests/tc_mode_check.py $tag 300 2>&1 | tail -3 | tee -a gpurun_out/tc_mode5.txt
done
AZ_TC_MODE=6 timeout -s KILL 90 python tests/tc_mode_check.py go9_c2 300 2>&1 | tail -3 | tee -a gpurun_out/tc_mode5.txt
date +%T
AZ_NODE_CACHE=0 AZ_COLLECT_OCC=0 timeout -s KILL 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r1c_replay.err | tee gpurun_out/bench_r1c_replay.json | cut -c1-300
date +%T
if grep -q "go9_c2.*OK" gpurun_out/tc_mode5.txt; then
  AZ_TC_MODE=5 timeout -s KILL 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r1c_mode5.err | tee gpurun_out/bench_r1c_mode5.json | cut -c1-300
fi
date +%T
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 140 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cold-start > gpurun_out/bench_under_ncu_r1c.log 2>&1
tail -2 gpurun_out/launches_r1c.csv | cut -c1-200
date +%T
