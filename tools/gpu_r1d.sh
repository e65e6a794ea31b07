#!/bin/bash
# Round-1 third session, second (last) GPU call: re-validate the full suite after the rt_alloc fix and the k_advance rework, calibrate the
# dense-x conv (mode 5) against mode 4 on realistic positions (multi-unit launches, four board geometries), bf16 tests under mode 5,
# benches of both, one full ncu capture of the dense-x kernel.
mkdir -p gpurun_out
date +%T
timeout -s KILL 300 python -m pytest tests -m gpu -q -n 4 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_r1d.txt
# anything that failed under four concurrent processes once more in a single process, with the assertion text
timeout -s KILL 200 python -m pytest tests -m gpu -q --lf --lfnf=none 2>&1 | tail -60 > gpurun_out/pytest_gpu_r1d_lf.txt; tail -3 gpurun_out/pytest_gpu_r1d_lf.txt
date +%T
for tag in go9_c2 gomoku13_c4 go9_small64 go19_128 go13_128; do
  cnt=1500; [ $tag = go19_128 ] && cnt=500
  timeout -s KILL 100 python tests/tc_mode_check.py $tag $cnt 4,5,6 2>&1 | grep -v Warning | tail -4 | tee -a gpurun_out/tc_modes_r1d.txt
done
date +%T
AZ_TC_MODE=5 timeout -s KILL 200 python -m pytest tests/test_gpu_engine.py::test_net_bf16_matches_bf16_emulation tests/test_gpu_facade.py tests/test_gpu_replay.py -q -n 4 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r1d_mode5.txt
date +%T
timeout -s KILL 200 python bench.py --steps 4 --warmup 3 2>gpurun_out/bench_r1d.err | tee gpurun_out/bench_r1d.json | cut -c1-300
tail -2 gpurun_out/bench_r1d.err
date +%T
AZ_TC_MODE=5 timeout -s KILL 200 python bench.py --steps 4 --warmup 3 2>gpurun_out/bench_r1d_mode5.err | tee gpurun_out/bench_r1d_mode5.json | cut -c1-300
tail -2 gpurun_out/bench_r1d_mode5.err
date +%T
AZ_TC_MODE=5 timeout -s KILL 150 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc_x -s 25 -c 2 -o gpurun_out/prof_dense_x -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cold-start --games 2048 > gpurun_out/ncu_dense_x.log 2>&1
tail -2 gpurun_out/ncu_dense_x.log
date +%T
