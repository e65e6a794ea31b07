#!/bin/bash
# round 2, last session: programmatic dependent launch on by default: full suite, smoke, default bench line
mkdir -p gpurun_out
timeout -s KILL 1200 python -X faulthandler -m pytest tests -m gpu -q -v > gpurun_out/pytest_gpu_r2p.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r2p.txt
timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-120
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_r2p.err | tee gpurun_out/bench_r2p.json | cut -c1-160
