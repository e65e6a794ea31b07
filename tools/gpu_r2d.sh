#!/bin/bash
# round 2, session d: batched-load tree_pick + redux argmax, fused tree pass (serial), persistent fused tree pass beside the conv kernel (AZ_PIPELINE=2)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2d.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2d.err | tee gpurun_out/bench_r2d.json | cut -c1-200
AZ_FUSED_TREE=0 timeout -s KILL 300 $B 2>gpurun_out/bench_r2d_unfused.err | tee gpurun_out/bench_r2d_unfused.json | cut -c1-200
AZ_PIPELINE=2 timeout -s KILL 300 $B 2>gpurun_out/bench_r2d_pipe2.err | tee gpurun_out/bench_r2d_pipe2.json | cut -c1-200
AZ_PIPELINE=1 timeout -s KILL 300 $B 2>gpurun_out/bench_r2d_pipe1.err | tee gpurun_out/bench_r2d_pipe1.json | cut -c1-200
timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2d_g512.err | tee gpurun_out/bench_r2d_g512.json | cut -c1-200
AZ_PIPELINE=2 timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2d_g512_pipe2.err | tee gpurun_out/bench_r2d_g512_pipe2.json | cut -c1-200
timeout -s KILL 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --weights ckpt 2>gpurun_out/bench_r2d_ckpt.err | tee gpurun_out/bench_r2d_ckpt.json | cut -c1-200
