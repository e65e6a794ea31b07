"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck): env steps, split-phase search, device self-play, both towers."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
sys.path.insert(0, 'tests')
from alpha_zero_b200.engine import Engine
from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
from fake_eval import make_fake_eval

torch.manual_seed(1)
for game, n, A, gomoku in (('go', 9, 82, False), ('gomoku', 9, 81, True)):
    net = randomize_batchnorm(AlphaZeroNet((17, n, n), A, 1, 64, 32, gomoku)).eval()
    for prec in ('fp32', 'bf16'):
        eng = Engine(game, n, num_games=6, max_simulations=24, max_parallel=4, net=(1, 64, 32), precision=prec, max_steps=14 if game == 'go' else 0)
        eng.set_weights(net.state_dict())
        rng = np.random.RandomState(0)
        for _ in range(10):
            lg = np.flatnonzero(eng.env_legal(0))
            eng.env_step([0], [int(rng.choice(lg))])
        ev = make_fake_eval(A)
        eng.search_begin([0, 2], [0, 0], 19652.0, 1.25, 16, 4, True, True, False)
        while True:
            obs, counts, active = eng.search_select()
            if active == 0:
                break
            if len(obs):
                p, v = ev(obs, True)
                eng.search_apply(np.stack(p), np.array(v, dtype=np.float32))
            else:
                eng.search_apply(None, None)
        r = eng.search_result(0)
        eng.search_commit(0, r['argmax'])
        eng.search_begin([1], [0], 19652.0, 1.25, 16, 4, False, False, True)
        eng.search_run()
        eng.selfplay_begin(16, 4, warm_up_steps=3, check_resign_after_steps=4, resign_threshold=-0.5, disable_resign_ratio=0.5)
        for _ in range(8):
            eng.selfplay_tick(10)
            eng.drain_games()
        c = eng.counters()
        print(game, prec, {k: c[k] for k in ('simulations', 'moves', 'games', 'errors')}, flush=True)
        eng.close()
print('sanity done')
