#!/bin/bash
# round 2, final validation of the committed build: full GPU suite, smoke, the default bench line (extras + CPU baseline), reference arm
mkdir -p gpurun_out
timeout -s KILL 1500 python -X faulthandler -m pytest tests -m gpu -q -v > gpurun_out/pytest_gpu_final.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.txt
timeout -s KILL 120 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke_final.txt
timeout -s KILL 600 python bench.py 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-200
AZ_REF_SECONDS=6 timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_final_ref.err | tee gpurun_out/bench_final_ref.json | cut -c1-200
