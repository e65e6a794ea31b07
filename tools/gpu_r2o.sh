#!/bin/bash
# round 2, session o: final build (device guard, tower ring) through the full suite; AZ_PDL=1 experiment
mkdir -p gpurun_out
timeout -s KILL 1500 python -X faulthandler -m pytest tests -m gpu -q -v > gpurun_out/pytest_gpu_r2o.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r2o.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2o.err | tee gpurun_out/bench_r2o.json | cut -c1-160
AZ_PDL=1 timeout -s KILL 300 $B 2>gpurun_out/bench_r2o_pdl.err | tee gpurun_out/bench_r2o_pdl.json | cut -c1-160
AZ_PDL=1 timeout -s KILL 600 python -m pytest tests/test_gpu_net_layers.py -m gpu -q -x 2>&1 | tail -2
AZ_PDL=1 timeout -s KILL 300 $B --games 512 2>gpurun_out/bench_r2o_pdl_g512.err | tee gpurun_out/bench_r2o_pdl_g512.json | cut -c1-160
