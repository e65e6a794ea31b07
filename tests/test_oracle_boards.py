"""Oracle board engines vs golden vectors recorded from the reference (CPU, no GPU needed)."""
import os

import numpy as np
import pytest

from oracle.boards import GoBoard, GomokuBoard
from trajectory import Trajectory, parse_corpus


def _replay(env, moves, keep=False):
    env.reset()
    tr = Trajectory(keep)
    played = 0
    for a in moves:
        if env.is_game_over():
            break
        _, reward, done, _ = env.step(int(a))
        tr.add(env.legal_actions, env.board, reward, done, env.to_play)
        played += 1
    return tr, played


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_selfplay_corpus_digests(golden_dir, game):
    """Every reference self-play SGF (2412 Go 9x9 + 2364 Gomoku 13x13): legal mask, board, reward, done, to_play per ply."""
    z = np.load(os.path.join(golden_dir, f'{game}_selfplay.npz'))
    games = parse_corpus(z)
    env = GoBoard(9, 7.5, 8) if game == 'go9' else GomokuBoard(13, 5, 8)
    stride = int(os.environ.get('AZ_CORPUS_STRIDE', '6'))  # full corpus with AZ_CORPUS_STRIDE=1
    bad = []
    for gi in range(0, len(games), stride):
        tr, played = _replay(env, games[gi])
        if tr.hexdigest() != str(z['digest'][gi]) or played != int(z['played'][gi]):
            bad.append(gi)
        else:
            res = env.get_result_string() if env.is_game_over() else ''
            assert res == str(z['result_env'][gi])
    assert not bad, f'{len(bad)} games differ, first {bad[:5]}'


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_selfplay_full_trajectories(golden_dir, game):
    z = np.load(os.path.join(golden_dir, f'{game}_selfplay.npz'))
    games = parse_corpus(z)
    env = GoBoard(9, 7.5, 8) if game == 'go9' else GomokuBoard(13, 5, 8)
    n_full = int(z['full_game'].max()) + 1
    row = 0
    for gi in range(n_full):
        tr, _ = _replay(env, games[gi], keep=True)
        for (legal, board, tail) in tr.rows:
            np.testing.assert_array_equal(legal, z['full_legal'][row])
            np.testing.assert_array_equal(board, z['full_board'][row])
            np.testing.assert_array_equal(tail, z['full_tail'][row])
            row += 1
        np.testing.assert_array_equal(env.observation(), z['full_obs_last'][gi])
    assert row == len(z['full_game'])


def _cases(z):
    return sorted({k.split('/')[0] for k in z.files if '/' in k})


def test_go19_unit_sequences(golden_dir):
    """The sequences pinned by the reference's unit tests (unit_tests/envs/go_test.py:80-276)."""
    z = np.load(os.path.join(golden_dir, 'go19_unit.npz'))
    for name in _cases(z):
        kw = {}
        if name.startswith('max_steps_'):
            kw['max_steps'] = int(name.split('_')[-1])
        if name == 'stacked_obs_4':
            kw['num_stack'] = 4
        env = GoBoard(19, 7.5, **kw)
        for a, r, d in zip(z[name + '/actions'], z[name + '/rewards'], z[name + '/dones']):
            _, reward, done, _ = env.step(int(a))
            assert reward == r and done == bool(d), name
        np.testing.assert_array_equal(np.asarray(env.legal_actions).astype(np.uint8), z[name + '/legal'], err_msg=name)
        np.testing.assert_array_equal(env.board.reshape(19, 19), z[name + '/board'], err_msg=name)
        np.testing.assert_array_equal(env.observation(), z[name + '/obs'], err_msg=name)
        assert (0 if env.winner is None else env.winner) == int(z[name + '/winner'][0]), name
        assert (env.get_result_string() if env.is_game_over() else '') == str(z[name + '/result'][0]), name
        if name + '/probe' in z.files:
            a = int(z[name + '/probe'][0])
            assert env.legal_actions[a] == 0 == int(z[name + '/probe_legal'][0])
            with pytest.raises(ValueError, match='Illegal action'):
                env.step(a)
    # the assertions go_test.py itself makes
    assert str(z['score_black/result'][0]).startswith('B+') and str(z['score_white/result'][0]).startswith('W+')
    env = GoBoard(19)
    with pytest.raises(ValueError, match='Invalid action'):
        env.step(500)
    for name in ('over_resign', 'over_pass'):
        env = GoBoard(19)
        for a in z[name + '/actions']:
            env.step(int(a))
        with pytest.raises(RuntimeError, match='Game is over'):
            env.step(6)


def test_gomoku_unit_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, 'gomoku_unit.npz'))
    for name in _cases(z):
        n, k = (int(v) for v in z[name + '/cfg'])
        env = GomokuBoard(n, k, 8)
        for a, r, d in zip(z[name + '/actions'], z[name + '/rewards'], z[name + '/dones']):
            _, reward, done, _ = env.step(int(a))
            assert reward == r and done == bool(d), name
        assert (0 if env.winner is None else env.winner) == int(z[name + '/winner'][0])
        assert env.get_result_string() == str(z[name + '/result'][0])
        np.testing.assert_array_equal(env.observation(), z[name + '/obs'])


EXTRA = [  # (file, constructor) -- random play and human games recorded from the reference (tests/golden/make_golden.py)
    ('random_go19', lambda: GoBoard(19, 7.5, 8)),
    ('random_go13', lambda: GoBoard(13, 5.5, 8, 300)),
    ('random_go9', lambda: GoBoard(9, 7.5, 8)),
    ('random_gomoku15', lambda: GomokuBoard(15, 5, 8)),
    ('pro_go9', lambda: GoBoard(9, 7.5, 8)),
]


@pytest.mark.parametrize('name', [e[0] for e in EXTRA])
def test_extra_corpora_digests(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f'{name}.npz'))
    games = parse_corpus(z)
    env = dict(EXTRA)[name]()
    stride = 1 if name != 'pro_go9' else int(os.environ.get('AZ_CORPUS_STRIDE', '6'))
    for gi in range(0, len(games), stride):
        tr, _ = _replay(env, games[gi])
        assert tr.hexdigest() == str(z['digest'][gi]), (name, gi)
        if 'result_env' in z.files:
            assert (env.get_result_string() if env.is_game_over() else '') == str(z['result_env'][gi])
