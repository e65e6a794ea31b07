#!/bin/bash
# multi-GPU session: N = $1 ranks.  Weak scaling (4096 games per GPU) with the device-packed NCCL sample gather, the strong split of
# ONE 4096-game population (BASELINE configs[2]), and the C4 / C5 shards (configs[3], [4]) when N = 8.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --no-extras"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout -s KILL 400 $RUN --steps 5 --warmup 3 2>gpurun_out/scale_weak_n$N.err | tail -1 | tee gpurun_out/scale_weak_n$N.json | cut -c1-250
timeout -s KILL 400 $RUN --steps 8 --warmup 3 --strong 2>gpurun_out/scale_strong_n$N.err | tail -1 | tee gpurun_out/scale_strong_n$N.json | cut -c1-250
if [ "$N" = "8" ]; then
  timeout -s KILL 400 $RUN --steps 8 --warmup 3 --workload gomoku13_c4 2>gpurun_out/scale_c4_n$N.err | tail -1 | tee gpurun_out/scale_c4_n$N.json | cut -c1-250
  timeout -s KILL 500 $RUN --steps 3 --warmup 3 --workload go19_c5 2>gpurun_out/scale_c5_n$N.err | tail -1 | tee gpurun_out/scale_c5_n$N.json | cut -c1-250
  timeout -s KILL 400 $RUN --steps 5 --warmup 3 --weights ckpt 2>gpurun_out/scale_ckpt_n$N.err | tail -1 | tee gpurun_out/scale_ckpt_n$N.json | cut -c1-250
fi
tail -3 gpurun_out/scale_weak_n$N.err
