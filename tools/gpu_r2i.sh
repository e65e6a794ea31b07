#!/bin/bash
# round 2, session i: terminal results cached in the node (select tail), full GPU suite on the final kernels
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r2i.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2i.err | tee gpurun_out/bench_r2i.json | cut -c1-200
timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2i_g512.err | tee gpurun_out/bench_r2i_g512.json | cut -c1-200
timeout -s KILL 300 $B --weights ckpt 2>gpurun_out/bench_r2i_ckpt.err | tee gpurun_out/bench_r2i_ckpt.json | cut -c1-200
