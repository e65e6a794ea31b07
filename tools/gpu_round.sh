#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests -m gpu -q 2>&1 | tail -${TAIL:-25} | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps ${STEPS:-3} --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ -n "$NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 140 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 25 -c 2 -o gpurun_out/prof_conv_tc -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --games 1024 > gpurun_out/ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_collect -s 60 -c 1 -o gpurun_out/prof_collect -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --games 1024 > gpurun_out/ncu_collect.log 2>&1
fi
