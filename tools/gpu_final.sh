#!/bin/bash
# Round-end evidence: other configs (sanity + numbers), final bench, ncu launch list + full captures.
mkdir -p gpurun_out
python bench.py --workload gomoku13_c4 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/c4.err | tee gpurun_out/bench_c4.json | cut -c1-300
tail -3 gpurun_out/c4.err
python bench.py --workload go19_c5 --steps 2 --warmup 3 --no-cpu-baseline 2>gpurun_out/c5.err | tee gpurun_out/bench_c5.json | cut -c1-300
tail -3 gpurun_out/c5.err
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_final.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 140 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_halo -s 25 -c 2 -o gpurun_out/prof_pair -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --games 2048 > gpurun_out/ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collect -s 60 -c 1 -o gpurun_out/prof_collect2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --games 4096 > gpurun_out/ncu_collect2.log 2>&1
