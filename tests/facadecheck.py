"""Checks of the reference-shaped Python API (envs / mcts / pipeline façades) against goldens recorded from the reference.
Run on the host-emulation binding (`-m "not gpu"`) and on the CUDA library (`-m gpu`)."""
import copy
import os
import pickle
import random

import numpy as np
import torch

from fake_eval import make_fake_eval

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def use_binding(binding):
    from alpha_zero_b200.envs import _pool

    _pool.reset_pools()
    _pool._TEST_BINDING = binding


def make_env(game, max_steps=None):
    from alpha_zero_b200.envs.go import GoEnv
    from alpha_zero_b200.envs.gomoku import GomokuEnv

    if game == 'go9':
        return GoEnv(komi=7.5, num_stack=8, board_size=9, **({} if max_steps is None else {'max_steps': max_steps}))
    return GomokuEnv(board_size=13, num_stack=8)


def env_contract():
    """The assertions of the reference's unit tests (unit_tests/envs/base_test.py, go_test.py, gomoku_test.py) on the façades."""
    import pytest

    from alpha_zero_b200.envs.go import GoEnv
    from alpha_zero_b200.envs.gomoku import GomokuEnv

    env = GoEnv(num_stack=8, board_size=19)
    obs = env.reset()
    assert env.action_space.n == 362 and env.observation_space.shape == (17, 19, 19) and obs.shape == (17, 19, 19)
    assert env.board.shape == (19, 19) and env.legal_actions.dtype == np.int64
    for gtp, want in (('A19', 0), ('T19', 18), ('A1', 19 * 18), ('T1', 360), ('C3', 19 * 16 + 2), ('E7', 19 * 12 + 4), ('PASS', 361)):
        assert env.gtp_to_action(gtp) == want  # go_test.py:48-61
    for bad in (500, 363, 999):
        with pytest.raises(ValueError, match='Invalid action'):
            env.step(bad)
    env.step(env.gtp_to_action('C7'))
    with pytest.raises(ValueError, match='Illegal action'):
        env.step(env.gtp_to_action('C7', check_illegal=False))
    assert env.steps == 1 and env.to_play == env.white_player and env.last_player == env.black_player
    # ko (go_test.py:114-127)
    env.reset()
    for b, w in zip(['A4', 'B4', 'C3', 'C1', 'D2'], ['A2', 'A3', 'B1', 'B3', 'C2']):
        env.step(env.gtp_to_action(b, check_illegal=False))
        env.step(env.gtp_to_action(w, check_illegal=False))
    env.step(env.gtp_to_action('B2'))
    with pytest.raises(ValueError, match='Illegal action'):
        env.step(env.gtp_to_action('C2', check_illegal=False))
    # resign / game over (go_test.py:129-139, 211-220)
    env.reset()
    for i in range(6):
        env.step(i)
    _, r, d, _ = env.step(env.resign_move)
    assert r == -1 and d and env.winner == env.white_player and env.get_result_string() == 'W+R'
    assert env.legal_actions.dtype == np.int8 and env.legal_actions.sum() == 0
    with pytest.raises(RuntimeError, match='Game is over'):
        env.step(6)
    # score known answer (go_test.py:175-209)
    env.reset()
    for g in ('C1', 'A1', 'B2', 'A2', 'A3', 'PASS', 'PASS'):
        _, r, d, _ = env.step(env.gtp_to_action(g))
    assert d and env.winner == env.black_player and r == 1.0 and env.get_result_string().startswith('B+')
    assert len(env.history) == 7 and env.history[0].color == 'B' and 'RE[B+' in env.to_sgf()
    # deepcopy and pickle keep the position; the copy evolves independently
    env.reset()
    for a in (3, 40, 41, 60):
        env.step(a)
    twin = copy.deepcopy(env)
    blob = pickle.dumps(env)
    twin.step(100)
    assert env.steps == 4 and twin.steps == 5
    back = pickle.loads(blob)
    np.testing.assert_array_equal(back.observation(), env.observation())
    np.testing.assert_array_equal(back.legal_actions, env.legal_actions)
    assert back.steps == 4 and [m.move for m in back.history] == [3, 40, 41, 60]
    # Gomoku (gomoku_test.py): five in a row for black on 7x7
    g = GomokuEnv(board_size=7, num_to_win=5, num_stack=8)
    g.reset()
    assert g.legal_actions.dtype == np.int8 and g.black_player == 1 and g.white_player == 2
    done = False
    for b, w in zip((0, 1, 2, 3, 4), (7, 8, 9, 10, 11)):
        _, r, done, _ = g.step(b)
        if done:
            break
        g.step(w)
    assert done and r == 1.0 and g.winner == 1 and g.get_result_string() == 'B+1.0' and g.board[0, 0] == 1 and g.board[1, 0] == 2


def mcts_api_traces(game):
    """uct_search / parallel_uct_search called exactly like the reference trace generator did: a seeded numpy RNG must give
    the same moves, the same pi (bit for bit: same numpy expressions on the same child_N) and the same Q values."""
    from alpha_zero_b200.mcts import parallel_uct_search, uct_search

    z = np.load(os.path.join(GOLDEN, f'mcts_{game}.npz'))
    A = 82 if game == 'go9' else 169
    ev = make_fake_eval(A)
    n_checked = 0
    for name in z['names']:
        prefix, plies, sims, par, noise, det, warm_steps, reuse, seed = (int(v) for v in z[f'{name}/cfg'])
        if sims > 100:
            continue
        np.random.seed(seed)
        env = make_env(game)
        env.reset()
        for a in z[f'{name}/prefix']:
            env.step(int(a))
        root = None
        for ply in range(len(z[f'{name}/move'])):
            warm = env.steps <= warm_steps
            if par > 1:
                mv, pi, rq, cq, nxt = parallel_uct_search(env, ev, root if reuse else None, 19652.0, 1.25, sims, par, bool(noise), warm, bool(det))
            else:
                mv, pi, rq, cq, nxt = uct_search(env, ev, root if reuse else None, 19652.0, 1.25, sims, bool(noise), warm, bool(det))
            tag = f'{game}/{name}/ply{ply}'
            assert mv == int(z[f'{name}/move'][ply]), tag
            assert pi.dtype == (np.float64 if game == 'go9' else np.float32), (tag, pi.dtype)
            np.testing.assert_array_equal(np.asarray(pi, dtype=np.float64), z[f'{name}/pi'][ply], err_msg=tag)
            assert float(rq) == z[f'{name}/root_q'][ply] and float(cq) == z[f'{name}/child_q'][ply], tag
            assert int(nxt is None) == int(z[f'{name}/next_is_none'][ply]), tag
            env.step(mv)
            root = nxt
            n_checked += 1
    return n_checked


def pipeline_traces(game):
    """play_and_record_one_game over the façade search with the oracle's torch-fp32 evaluator and the reference's seeds:
    same move history, pi, z and stats dict as the reference run."""
    from alpha_zero_b200.mcts import parallel_uct_search, uct_search
    from alpha_zero_b200.pipeline import play_and_record_one_game
    from oracle import net as onet

    z = np.load(os.path.join(GOLDEN, f'pipeline_{game}.npz'))
    zn = np.load(os.path.join(GOLDEN, 'net.npz'))
    tag = f'{game}_small'
    sd = {k[len(tag) + 4:]: torch.from_numpy(zn[k]) for k in zn.files if k.startswith(tag + '/sd/')}
    torch.set_num_threads(1)
    ev = onet.make_eval_func(sd, gomoku=(game != 'go9'))
    for name in sorted({k.split('/')[0] for k in z.files if '/' in k}):
        nb, nf, fc, sims, par, warm, chk, resign_disabled, max_steps, seed = (int(v) for v in z[f'{name}/cfg'])
        thr = float(z[f'{name}/thr'][0])
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        env = make_env(game, max_steps)

        def player(env, root_node, c_puct_base, c_puct_init, warm_up=False):
            if par > 1:
                return parallel_uct_search(env, ev, root_node, c_puct_base, c_puct_init, sims, par, True, warm_up, False)
            return uct_search(env, ev, root_node, c_puct_base, c_puct_init, sims, True, warm_up, False)

        seq, stats = play_and_record_one_game(env, player, bool(resign_disabled), 19652.0, 1.25, warm, chk, thr, None)
        np.testing.assert_array_equal(np.array([m.move for m in env.history], dtype=np.int32), z[f'{name}/history'], err_msg=name)
        assert repr({k: stats[k] for k in sorted(stats)}) == str(z[f'{name}/stats_repr'][0]), (stats, str(z[f'{name}/stats_repr'][0]))
        np.testing.assert_array_equal(np.stack([t.state for t in seq]), z[f'{name}/states'])
        np.testing.assert_allclose(np.stack([np.asarray(t.pi_prob, dtype=np.float64) for t in seq]), z[f'{name}/pis'], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(np.array([t.value for t in seq], dtype=np.float32), z[f'{name}/values'])


def error_paths():
    """Argument / state errors surface as the reference's exception types (mcts_v2.py:356-361) or as EngineError."""
    import pytest

    from alpha_zero_b200._lib import EngineError
    from alpha_zero_b200.engine import Engine
    from alpha_zero_b200.envs import _pool
    from alpha_zero_b200.mcts import Node, parallel_uct_search, uct_search

    ev = make_fake_eval(82)
    env = make_env('go9')
    env.reset()
    with pytest.raises(ValueError, match='num_simulations'):
        uct_search(env, ev, None, 19652.0, 1.25, 0)
    with pytest.raises(ValueError, match='BoardGameEnv'):
        uct_search(object(), ev, None, 19652.0, 1.25, 8)
    with pytest.raises(ValueError, match='root_node'):
        parallel_uct_search(env, ev, Node(to_play=1, num_actions=82), 19652.0, 1.25, 8, 4)  # a handle that is not this env's tree
    mv, pi, rq, cq, nxt = parallel_uct_search(env, ev, None, 19652.0, 1.25, 16, 4, False, True, True)
    env.step(mv)
    other = make_env('go9')
    other.reset()
    with pytest.raises(ValueError, match='root_node'):
        uct_search(other, ev, nxt, 19652.0, 1.25, 8)  # subtree of another env
    env.step(env.pass_move)
    env.step(env.pass_move)
    assert env.is_game_over()
    with pytest.raises(RuntimeError, match='Game is over'):
        uct_search(env, ev, None, 19652.0, 1.25, 8)
    # engine-level capacity and argument checks
    eng = Engine('go', 9, num_games=2, max_simulations=32, max_parallel=4, binding=_pool._TEST_BINDING)
    with pytest.raises(EngineError, match='node pool'):
        eng.search_begin([0], [0], 19652.0, 1.25, 4000, 4)
    with pytest.raises(EngineError, match='max_parallel'):
        eng.search_begin([0], [0], 19652.0, 1.25, 16, 64)
    with pytest.raises(ValueError, match='slot'):
        eng.env_reset([5])
    with pytest.raises(ValueError, match='ascending'):
        eng.search_begin([1, 0], [0, 0], 19652.0, 1.25, 16, 4)
    with pytest.raises(EngineError, match='no search in progress|not finished'):
        eng.search_result(1)
    with pytest.raises(Exception):
        Engine('go', 25, num_games=1, binding=_pool._TEST_BINDING)  # board size out of range
    eng.close()


def batched_matches(game):
    """play_matches (evaluation matches, many games in lock step) vs the oracle playing the same match alone: with one game and
    a seeded numpy RNG the move list is identical (no reuse, no noise, T=0.1 sampling); with several games every game is a legal,
    finished game whose result string matches an oracle replay."""
    from alpha_zero_b200.envs import _pool
    from alpha_zero_b200.matches import play_matches
    from oracle.boards import GoBoard, GomokuBoard
    from oracle.search import search

    kind, n, A = ('go', 9, 82) if game == 'go9' else ('gomoku', 13, 169)
    ev_b, ev_w = make_fake_eval(A, sharpness=2.0), make_fake_eval(A, sharpness=3.0)
    kw = dict(num_simulations=24, num_parallel=4, max_steps=30 if kind == 'go' else 0, binding=_pool._TEST_BINDING)
    np.random.seed(21)
    one = play_matches(1, kind, n, ev_b, ev_w, **kw)[0]
    np.random.seed(21)
    env = GoBoard(9, 7.5, 8, 30) if kind == 'go' else GomokuBoard(13, 5, 8)
    mvs, side = [], 0
    while not env.is_game_over():
        mv, *_ = search(env, (ev_b, ev_w)[side], None, 19652.0, 1.25, 24, 4, False, False, False)
        env.step(mv)
        mvs.append(mv)
        side ^= 1
    assert one['moves'] == mvs and one['game_result'] == env.get_result_string() and one['game_length'] == env.steps
    np.random.seed(22)
    many = play_matches(6, kind, n, ev_b, ev_w, **kw)
    assert len({tuple(g['moves']) for g in many}) > 1  # sampled games differ
    for g in many:
        env.reset()
        for mv in g['moves']:
            assert env.legal_actions[mv] == 1
            env.step(mv)
        assert env.is_game_over() and env.get_result_string() == g['game_result'] and env.steps == g['game_length']


def device_matches(game, binding=None, on_gpu=False):
    """play_matches_on_device (both weight sets in one engine, the loop on the device: az_match_begin / az_match_tick).
    Every game is a legal, finished game whose result string equals an oracle replay; games are numbered and coloured as
    documented (odd games swapped); deterministic games repeat exactly.  On the GPU the deterministic match must also equal the
    host-driven play_matches (split-phase search + az_net_forward of each network) move for move, in both colour assignments."""
    import torch

    from alpha_zero_b200.matches import play_matches, play_matches_on_device
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from oracle.boards import GoBoard, GomokuBoard

    kind, n, A = ('go', 9, 82) if game == 'go9' else ('gomoku', 13, 169)
    dims = (2, 64, 64) if on_gpu else (1, 16, 16)
    nets = []
    for seed in (11, 12):
        torch.manual_seed(seed)
        nets.append(randomize_batchnorm(AlphaZeroNet((17, n, n), A, *dims, kind == 'gomoku'), seed=seed).eval())
    kw = dict(num_simulations=24, num_parallel=4, max_steps=30 if kind == 'go' else 0, binding=binding, precision='fp32')

    def replay(g):
        env = GoBoard(9, 7.5, 8, 30) if kind == 'go' else GomokuBoard(13, 5, 8)
        for mv in g['moves']:
            assert env.legal_actions[mv] == 1
            env.step(mv)
        assert env.is_game_over() and env.get_result_string() == g['game_result'] and env.steps == g['game_length']

    # 7 games on 3 slots: three rounds per slot, odd slot count -> colours alternate inside a slot
    many = play_matches_on_device(7, kind, n, nets[0], nets[1], swap_colours=True, slots=3, seed=5, **kw)
    assert [g['game'] for g in many] == list(range(7))
    assert [g['black_is'] for g in many] == ['black', 'white'] * 3 + ['black']
    for g in many:
        replay(g)
    assert len({tuple(g['moves']) for g in many}) > 1  # sampled games differ
    assert many[0]['engine_counters']['simulations'] > 0
    det = play_matches_on_device(4, kind, n, nets[0], nets[1], swap_colours=True, slots=4, deterministic=True, **kw)
    assert det[0]['moves'] == det[2]['moves'] and det[1]['moves'] == det[3]['moves']
    for g in det:
        replay(g)
    if on_gpu:
        os.environ['AZ_NET_PRECISION'] = 'fp32'
        try:
            host = play_matches(1, kind, n, nets[0], nets[1], num_simulations=24, num_parallel=4, max_steps=30 if kind == 'go' else 0, deterministic=True)
            host_swapped = play_matches(1, kind, n, nets[1], nets[0], num_simulations=24, num_parallel=4, max_steps=30 if kind == 'go' else 0, deterministic=True)
        finally:
            os.environ.pop('AZ_NET_PRECISION', None)
        assert det[0]['moves'] == host[0]['moves'] and det[0]['game_result'] == host[0]['game_result']
        assert det[1]['moves'] == host_swapped[0]['moves'] and det[1]['game_result'] == host_swapped[0]['game_result']
