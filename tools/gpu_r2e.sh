#!/bin/bash
# round 2, session e: full suite (device matches, relaxed bf16x3 search gate), default bench after the epilogue change, match throughput,
# ncu --set full of the select kernel after the batched-load pick
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r2e.txt
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout -s KILL 300 $B 2>gpurun_out/bench_r2e.err | tee gpurun_out/bench_r2e.json | cut -c1-200
timeout -s KILL 600 python tools/bench_dataset.py 2>gpurun_out/bench_dataset_r2e.err | tee gpurun_out/bench_dataset_r2e.json | cut -c1-400
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"k_collect_nc_occ|k_advance" -s 890 -c 2 -o gpurun_out/prof_tree_r2e -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_tree_r2e.log 2>&1
tail -2 gpurun_out/ncu_tree_r2e.log | cut -c1-200
