"""Throughput of the learner input path: minibatch gather + dihedral transformation from the device
replay into CUDA tensors, vs the reference's host path (numpy stack of Transitions + torchvision transform on the GPU)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
from alpha_zero_b200.engine import Engine

n, A, cap, B = 9, 82, 500_000, 4096
eng = Engine('go', n, num_games=1, max_simulations=8, max_parallel=1)
eng.replay_create(cap)
rng = np.random.RandomState(0)
chunk = 50_000
st = (rng.rand(chunk, 17, n, n) < 0.3).astype(np.int8)
pi = rng.rand(chunk, A).astype(np.float32)
z = rng.choice([-1.0, 1.0], size=chunk).astype(np.float32)
for _ in range(cap // chunk):
    eng.replay_add(st, pi, z)
ts = torch.empty((B, 17, n, n), dtype=torch.int8, device='cuda')
tp = torch.empty((B, A), dtype=torch.float32, device='cuda')
tz = torch.empty((B,), dtype=torch.float32, device='cuda')
out = (ts.data_ptr(), tp.data_ptr(), tz.data_ptr())
stream = torch.cuda.ExternalStream(eng.stream())
for t in (0, 3):
    idx = rng.randint(0, cap, size=B).astype(np.int32)
    for _ in range(5):
        eng.replay_sample(idx, t, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iters = 200
    for _ in range(iters):
        idx = rng.randint(0, cap, size=B).astype(np.int32)
        eng.replay_sample(idx, t, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    bytes_moved = B * (17 * n * n + 4 * A + 4) * 2  # read + write
    print(json.dumps({'metric': 'replay_minibatch_samples_per_sec', 'transform': t, 'batch': B, 'value': B / dt, 'ms_per_batch': dt * 1e3,
                      'achieved_GBps': bytes_moved / dt / 1e9, 'note': 'end to end through az_replay_sample (index H2D + gather kernel + sync)'}))
eng.close()
