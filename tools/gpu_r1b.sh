#!/bin/bash
# Late round-1 validation: full GPU suite (incl. matches / dataset on-ramp), on-ramp throughput, final bench, fresh launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r1b.txt
timeout 400 python tools/bench_dataset.py 2>gpurun_out/bench_dataset.err | tee gpurun_out/bench_dataset.json | cut -c1-400
tail -3 gpurun_out/bench_dataset.err
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_r1b.json | cut -c1-400
tail -2 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 140 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_r1b.log 2>&1
tail -2 gpurun_out/launches_r1b.csv | cut -c1-200
