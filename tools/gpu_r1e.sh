#!/bin/bash
# Round-1 third session, last GPU call (4 minutes left): the full suite and the bench under the final defaults
# (dense-x tower, node-cache search, two-pass re-root).
mkdir -p gpurun_out
date +%T
timeout -s KILL 170 python -m pytest tests -m gpu -q -n 6 2>&1 | tail -40 > gpurun_out/pytest_gpu_r1e.txt; tail -12 gpurun_out/pytest_gpu_r1e.txt
date +%T
timeout -s KILL 100 python bench.py --steps 4 --warmup 3 2>gpurun_out/bench_r1e.err | tee gpurun_out/bench_r1e.json | cut -c1-300
tail -2 gpurun_out/bench_r1e.err
date +%T
