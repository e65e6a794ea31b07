#!/bin/bash
# round 2, session f: full GPU suite with complete log (a crash in session e), deferred re-root payload, C4 with the halo kernel
mkdir -p gpurun_out
timeout -s KILL 1500 python -X faulthandler -m pytest tests -m gpu -q -v > gpurun_out/pytest_gpu_r2f_full.txt 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|Fatal|Segmentation|Aborted" gpurun_out/pytest_gpu_r2f_full.txt | tail -8
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2f.err | tee gpurun_out/bench_r2f.json | cut -c1-200
AZ_DEFER_REROOT=0 timeout -s KILL 300 $B 2>gpurun_out/bench_r2f_nodefer.err | tee gpurun_out/bench_r2f_nodefer.json | cut -c1-200
timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2f_g512.err | tee gpurun_out/bench_r2f_g512.json | cut -c1-200
timeout -s KILL 300 $B --workload gomoku13_c4 2>gpurun_out/bench_r2f_c4.err | tee gpurun_out/bench_r2f_c4.json | cut -c1-200
AZ_TC_MODE=4 timeout -s KILL 300 $B --workload gomoku13_c4 2>gpurun_out/bench_r2f_c4_mode4.err | tee gpurun_out/bench_r2f_c4_mode4.json | cut -c1-200
timeout -s KILL 300 $B --workload go19_c5 --steps 2 2>gpurun_out/bench_r2f_c5.err | tee gpurun_out/bench_r2f_c5.json | cut -c1-200
