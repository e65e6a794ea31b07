#!/bin/bash
# round 2, session k: ncu --set full of the small kernels of the tick (cold-start population: no 384-tick prologue to skip under the
# profiler), and a 66-step (~60 s) window for the windowed games/s
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none -k regex:"k_heads|k_net_input|k_apply|k_compact|k_reroot_payload|k_advance_d" -s 366 -c 6 -o gpurun_out/prof_aux_r2k -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --cold-start > gpurun_out/ncu_aux_r2k.log 2>&1
tail -1 gpurun_out/ncu_aux_r2k.log | cut -c1-120; ls -la gpurun_out | head
timeout -s KILL 400 python bench.py --steps 66 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/bench_r2k_60s.err | tee gpurun_out/bench_r2k_60s.json | cut -c1-200
