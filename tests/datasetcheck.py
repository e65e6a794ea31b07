"""Checks of the evaluation-dataset on-ramp (alpha_zero_b200/eval_dataset.py, az_env_replay) shared by the emulation test
(tests/test_emu_dataset.py) and the GPU test (tests/test_gpu_dataset.py).  Golden: tests/golden/eval_dataset_go9.npz = the
reference's own replay_sgf on 410 recorded games (make_golden.py part_eval_dataset_go9)."""
import hashlib
import os

import numpy as np

from alpha_zero_b200 import eval_dataset as ed
from alpha_zero_b200.engine import Engine

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'eval_dataset_go9.npz')


def replay_matches_reference(binding):
    g = np.load(GOLDEN)
    ed.reset_filters()
    games = [(str(n), str(t)) for n, t in zip(g['names'], g['texts'])]
    # small waves so that several kernel launches and slot reuse are exercised
    hist = ed.replay_sgf_games(games, 8, board_size=9, binding=binding, max_slots=96, max_positions=6000)
    assert len(hist) == len(games)
    for i, h in enumerate(hist):
        assert (h is not None) == bool(g['valid'][i]), (games[i][0], 'kept' if h is not None else 'dropped')
        if h is None:
            continue
        states, moves, values = h
        assert len(states) == len(moves) == len(values) == int(g['counts'][i]), games[i][0]
        sha = hashlib.sha1()
        for t in range(len(moves)):
            sha.update(states[t].tobytes())
            sha.update(np.int32(moves[t]).tobytes())
            sha.update(np.float32(values[t]).tobytes())
        assert sha.hexdigest() == str(g['digest'][i]), games[i][0]
    want = dict(zip((str(k) for k in g['mismatch_keys']), (int(v) for v in g['mismatch_values'])))
    assert ed.MISMATCH_GAMES == want
    ed.reset_filters()


def replay_abi_edges(binding):
    """az_env_replay stops where step() would raise, keeps the final position in the slot, and validates its arguments."""
    import pytest

    eng = Engine('go', 9, num_games=4, max_simulations=1, max_parallel=1, komi=7.5, net=None, binding=binding)
    # game 0: fine; game 1: occupied point at move 2; game 2: two passes then a move (game over); game 3: empty list
    games = [[40, 41, 81, 30], [10, 11, 10, 12], [81, 81, 5], []]
    states, off, played, status = eng.env_replay([0, 1, 2, 3], games)
    assert list(off) == [0, 4, 8, 11, 11]
    assert list(played) == [4, 2, 2, 0] and list(status) == [0, -3, -4, 0]
    assert states.shape == (11, 17, 9, 9)
    assert not states[0][:16].any() and states[0][16].all()  # empty board, black to play (envs/base.py:228-259)
    # slot 0 against a step-by-step replay on another slot
    eng.env_reset([3])
    for t, a in enumerate(games[0]):
        assert np.array_equal(eng.env_observation(3), states[t]), t
        eng.env_step([3], [a])
    assert np.array_equal(eng.env_board(3), eng.env_board(0)) and eng.env_scalars(3) == eng.env_scalars(0)
    assert not states[off[1] + 2:off[2]].any()  # rows after the rejected move stay zero
    assert eng.env_scalars(1)['steps'] == 2 and eng.env_scalars(2)['done'] == 1
    _, _, played, status = eng.env_replay([2], [[0, 1, 2]], want_states=False)  # the slot is reset first
    assert list(played) == [3] and list(status) == [0]
    with pytest.raises(ValueError):
        eng.env_replay([0, 0], [[1], [2]])
    with pytest.raises(ValueError):
        eng.env_replay([7], [[1]])
    eng.close()


def dataset_and_metrics(binding, tmp_path, net_factory=None):
    """build_eval_dataset on a directory of SGF files == concatenation of the reference's kept games (tensor shapes / dtypes of
    core/eval_dataset.py:262-266)."""
    import torch

    g = np.load(GOLDEN)
    ed.reset_filters()
    for i in range(40):
        with open(os.path.join(tmp_path, f'{i:03d}.sgf'), 'w') as f:
            f.write(str(g['texts'][i]))
    ds = ed.build_eval_dataset(str(tmp_path), 8, board_size=9, binding=binding)
    states, target_pi, target_v = ds.tensors
    assert states.dtype == torch.float32 and target_pi.dtype == torch.float32 and target_v.dtype == torch.float32
    assert states.shape[1:] == (17, 9, 9) and target_pi.shape[1] == 82
    assert torch.all(target_pi.sum(dim=1) == 1.0)
    # os.walk order is arbitrary, so compare as a multiset of per-game position counts (no duplicates among these 40 files)
    assert len(ds) == int(np.sum(np.maximum(g['counts'][:40], 0)))
    ed.reset_filters()
    return ds
