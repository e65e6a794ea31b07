"""Interop with the UNMODIFIED reference learner side (only where /root/reference exists, i.e. in the build container):
the queue items our actor emits must be consumable by the reference's UniformReplay and compute_losses unchanged."""
import os
import sys

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference checkout not present (GPU box)')


def _fake_finished_games(n_games, A, n, has_pass):
    rng = np.random.RandomState(0)
    games, states, pis, zs, moves = [], [], [], [], []
    for g in range(n_games):
        ln = int(rng.randint(5, 12))
        games.append(dict(slot=g, game_length=ln, winner=1 if g % 2 == 0 else (-1 if has_pass else 2), by_resign=int(has_pass and g % 3 == 0), score=3.5,
                          num_passes=1, is_resign_disabled=1, is_marked_for_resign=0, is_could_won=0, marked_resign_player=0,
                          first_sample=len(zs), reserved=g))
        states.append((rng.rand(ln, 17, n, n) < 0.2).astype(np.int8))
        p = rng.rand(ln, A).astype(np.float32)
        pis.append(p / p.sum(axis=1, keepdims=True))
        zs.append(rng.choice([-1.0, 1.0], size=ln).astype(np.float32))
        moves.append(rng.randint(0, A, size=ln).astype(np.int16))
    return games, np.concatenate(states), np.concatenate(pis), np.concatenate(zs), np.concatenate(moves)


@pytest.mark.parametrize('game', ['go', 'gomoku'])
def test_items_feed_reference_replay_and_losses(game):
    for p in (os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault('BOARD_SIZE', '9')
    import torch
    from alpha_zero.core.network import AlphaZeroNet as RefNet
    from alpha_zero.core.pipeline import compute_losses
    from alpha_zero.core.replay import Transition as RefTransition, UniformReplay

    from alpha_zero_b200.envs.go import GoEnv
    from alpha_zero_b200.envs.gomoku import GomokuEnv
    from alpha_zero_b200.pipeline import games_to_queue_items
    from alpha_zero_b200.replay import Transition

    assert Transition._fields == RefTransition._fields  # same record; the reference's own class is used when importable at import time
    n = 9
    env = GoEnv(board_size=n) if game == 'go' else GomokuEnv(board_size=n)
    A = env.action_dim
    games, states, pis, zs, moves = _fake_finished_games(6, A, n, env.has_pass_move)
    items = games_to_queue_items(None, env, games, states, pis, zs, moves, resign_threshold=-0.9)
    replay = UniformReplay(capacity=1000, random_state=np.random.RandomState(1), compress_data=False)
    for seq, stats, history in items:
        assert stats['game_length'] == len(seq) and isinstance(stats['game_result'], str)
        assert ('num_passes' in stats) == env.has_pass_move and ('is_resign_disabled' in stats) == env.has_resign_move
        assert len(history) <= len(seq) and all(c in ('B', 'W') for c, _ in history)
        replay.add_game(seq)
    assert replay.num_games_added == 6 and replay.size == sum(g['game_length'] for g in games)
    batch = replay.sample(16)
    assert batch.state.shape == (16, 17, n, n) and batch.pi_prob.shape == (16, A) and batch.value.shape == (16,)
    net = RefNet((17, n, n), A, 1, 16, 16, game == 'gomoku')
    policy_loss, value_loss = compute_losses(net, torch.device('cpu'), batch, False)
    assert torch.isfinite(policy_loss) and torch.isfinite(value_loss)
    (policy_loss + value_loss).backward()


def test_actor_accepts_reference_env_objects():
    """run_selfplay_actor_loop only reads settings from `env`: the reference's GoEnv / GomokuEnv instances map to the right engine."""
    for p in (os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault('BOARD_SIZE', '9')
    from alpha_zero.envs.go import GoEnv as RefGo
    from alpha_zero.envs.gomoku import GomokuEnv as RefGomoku

    from alpha_zero_b200.pipeline import _env_kind, games_to_queue_items

    go, gm = RefGo(komi=5.5, num_stack=8, max_steps=60), RefGomoku(board_size=13, num_to_win=5, num_stack=8)
    assert _env_kind(go) == 'go' and _env_kind(gm) == 'gomoku'
    assert (go.board_size, go.komi, go.max_steps, go.num_stack) == (9, 5.5, 60, 8) and gm.num_to_win == 5
    games, states, pis, zs, moves = _fake_finished_games(2, 82, 9, True)
    (seq, stats, hist), _ = games_to_queue_items(None, go, games, states, pis, zs, moves, resign_threshold=-0.9)
    assert stats['game_result'] in ('B+R', 'W+R', 'B+3.5') and stats['marked_resign_player'] is None and 'num_passes' in stats
