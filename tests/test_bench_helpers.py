"""bench.py's checker pieces on the host-emulation engine: the oracle replay of drained games accepts what the engine produced
and rejects a tampered record; the strong-scaling split covers every game exactly once."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402
from alpha_zero_b200.engine import Engine  # noqa: E402
from alpha_zero_b200.gather import shard_slots  # noqa: E402
from test_emu_selfplay import _dummy_weights  # noqa: E402

import bench  # noqa: E402


def test_oracle_replay_accepts_engine_games_and_rejects_tampering():
    emu = Binding(ctypes.CDLL(build_emu.build()))
    eng = Engine('go', 9, num_games=12, max_simulations=16, max_parallel=4, net=(1, 16, 16), precision='fp32', seed=3, sample_ring=4000, binding=emu)
    eng.set_weights(_dummy_weights(1, 16, 16, 17, 82, 81))
    eng.selfplay_begin(12, 4, warm_up_steps=6, check_resign_after_steps=8, resign_threshold=-0.55, disable_resign_ratio=0.5)
    games = []
    for _ in range(80):
        eng.selfplay_tick(8)
        games, st, pis, zs = eng.drain_games(copy=True)
        if len(games) >= 3:
            break
    assert len(games) >= 3
    mv = eng.last_moves.copy()
    res = bench.oracle_replay('go', 9, games, st, pis, zs, mv, limit=8)
    assert res['ok'] and res['games'] == min(8, len(games))
    bad_z = zs.copy()
    g = games[0]
    flip = g['first_sample'] + g['game_length'] // 2
    bad_z[flip] = -bad_z[flip] if bad_z[flip] != 0 else 1.0
    assert not bench.oracle_replay('go', 9, games, st, pis, bad_z, mv, limit=8)['ok']
    bad_st = st.copy()
    bad_st[g['first_sample'] + 1, 0, 4, 4] ^= 1
    assert not bench.oracle_replay('go', 9, games, bad_st, pis, zs, mv, limit=8)['ok']
    eng.close()


def test_strong_split_covers_every_game_once():
    for total, world in ((4096, 8), (4096, 3), (1024, 8), (10, 4)):
        spans = [shard_slots(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
