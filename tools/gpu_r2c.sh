#!/bin/bash
# round 2, session c: full GPU suite on the new build, the new bench line (extras + one CPU procedure), ncu --set full of the tower
# kernel (plain + residual launch) and of the tree kernels on the age-staggered population, per-GPU shards of the strong-scaling split
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2c.txt
timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_r2c.txt
timeout -s KILL 400 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_r2c.err | tee gpurun_out/bench_r2c.json | cut -c1-300
for g in 2048 1024 512; do
  timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --games $g 2>gpurun_out/bench_r2c_g$g.err | tee gpurun_out/bench_r2c_g$g.json | cut -c1-200
done
# staggered population under ncu: the prologue (128 x 3 ticks) and 3 warm-up steps are skipped by launch count, only the named kernel is replayed
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_x -s 9241 -c 2 -o gpurun_out/prof_conv_r2c -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_conv_r2c.log 2>&1
tail -2 gpurun_out/ncu_conv_r2c.log | cut -c1-200
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"k_collect_nc_occ|k_advance|k_apply" -s 1335 -c 3 -o gpurun_out/prof_tree_r2c -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_tree_r2c.log 2>&1
tail -2 gpurun_out/ncu_tree_r2c.log | cut -c1-200
