"""Reference-shaped Python API on the CUDA library (B200): the same checks as tests/test_emu_facade.py, plus
create_mcts_player with the CUDA network and the batched actor loop."""
import multiprocessing as mp
import os
import queue
import threading

import numpy as np
import pytest
import torch

import facadecheck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def cuda_binding():
    facadecheck.use_binding(None)  # product path: alpha_zero_b200._lib.load()
    yield
    facadecheck.use_binding(None)


def test_env_contract():
    facadecheck.env_contract()


def test_error_paths():
    facadecheck.error_paths()


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_mcts_api_traces(game):
    assert facadecheck.mcts_api_traces(game) > 40


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_pipeline_traces(game):
    facadecheck.pipeline_traces(game)


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_batched_matches(game):
    facadecheck.batched_matches(game)


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_device_matches(game):
    facadecheck.device_matches(game, on_gpu=True)


def test_create_mcts_player_matches_oracle_game():
    """create_mcts_player(network, device, ...) + play_and_record_one_game with the CUDA fp32 tower vs the oracle game loop with
    the torch fp32 net under the same numpy seed: identical move history and z, pi within 1e-3."""
    from alpha_zero_b200.envs.go import GoEnv
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from alpha_zero_b200.pipeline import create_mcts_player, play_and_record_one_game
    from oracle import net as onet
    from oracle.boards import GoBoard
    from oracle.selfplay import play_one_game

    torch.manual_seed(123)
    net = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 82, 2, 32, 32, False)).eval()
    np.random.seed(5)
    env = GoEnv(komi=7.5, num_stack=8, max_steps=40, board_size=9)
    player = create_mcts_player(net, torch.device('cuda:0'), 24, 4, root_noise=True, deterministic=False)
    seq, stats = play_and_record_one_game(env, player, True, 19652.0, 1.25, 6, 4, -1.0, None)
    np.random.seed(5)
    oenv = GoBoard(9, 7.5, 8, 40)
    oseq, ostats = play_one_game(oenv, onet.make_eval_func(net.state_dict(), False), 24, 4, True, 19652.0, 1.25, 6, 4, -1.0)
    assert [m.move for m in env.history] == list(oenv.history)
    assert stats == ostats
    np.testing.assert_allclose(np.stack([t.pi_prob for t in seq]), np.stack([p for _, p, _ in oseq]), rtol=0, atol=1e-3)
    assert [t.value for t in seq] == [v for _, _, v in oseq]


def test_actor_loop_emits_reference_shaped_items(tmp_path):
    from alpha_zero_b200.envs.go import GoEnv
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from alpha_zero_b200.pipeline import run_selfplay_actor_loop

    os.environ['AZ_ACTOR_GAMES'] = '64'
    os.environ['AZ_NET_PRECISION'] = 'bf16'
    torch.manual_seed(123)
    net = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 82, 2, 64, 64, False)).eval()
    env = GoEnv(komi=7.5, num_stack=8, max_steps=24, board_size=9)
    q = queue.Queue()
    stop, ckpt = threading.Event(), threading.Event()
    var_ckpt = mp.Value('c', b'') if False else type('V', (), {'value': b''})()
    var_thr = type('V', (), {'value': -0.9})()
    t = threading.Thread(target=run_selfplay_actor_loop, args=(1, 0, net, torch.device('cuda:0'), q, env, 16, 4, 19652.0, 1.25, 4, 6, 0.5, str(tmp_path), 5,
                                                               str(tmp_path), None, 'INFO', var_ckpt, var_thr, ckpt, stop))
    t.start()
    items = []
    try:
        while len(items) < 70:
            items.append(q.get(timeout=120))
    finally:
        stop.set()
        t.join(timeout=120)
    for seq, stats in items:
        assert stats['game_length'] == len(seq) and seq[0].state.shape == (17, 9, 9) and seq[0].pi_prob.shape == (82,)
        assert set(stats) >= {'game_length', 'game_result', 'num_passes', 'is_resign_disabled', 'is_marked_for_resign', 'is_could_won',
                              'marked_resign_player', 'resign_threshold', 'time_per_game', 'training_steps'}
        assert abs(float(seq[0].pi_prob.sum()) - 1.0) < 1e-5 and seq[0].value in (-1.0, 0.0, 1.0)
    assert os.path.exists(tmp_path / 'actor0.csv') and any(f.endswith('.sgf') for f in os.listdir(tmp_path))


def test_actor_loop_as_spawned_process_with_ckpt_hot_swap(tmp_path):
    """The deployment shape of the reference (training_go.py:276-345, 393): `spawn` start method, the env and the network
    pickled into the child, manager Values for the checkpoint path / resign threshold, mp.Events, a bounded mp.Queue.
    After the learner 'publishes' a checkpoint the emitted games carry its training_steps."""
    from alpha_zero_b200.envs.gomoku import GomokuEnv
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from alpha_zero_b200.pipeline import run_selfplay_actor_loop

    os.environ['AZ_ACTOR_GAMES'] = '64'
    os.environ['AZ_NET_PRECISION'] = 'bf16'
    ctx = mp.get_context('spawn')
    torch.manual_seed(123)
    net = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 81, 2, 64, 64, True)).eval()
    env = GomokuEnv(board_size=9, num_stack=8)
    ckpt = tmp_path / 'training_steps_7.ckpt'
    torch.save({'network': net.state_dict(), 'training_steps': 7}, ckpt)
    stop, pause = ctx.Event(), ctx.Event()
    q = ctx.Queue(maxsize=4)
    with ctx.Manager() as manager:
        var_ckpt = manager.Value('s', b'')
        var_thr = manager.Value('d', -1.0)
        p = ctx.Process(target=run_selfplay_actor_loop, args=(1, 3, net, torch.device('cuda:0'), q, env, 16, 4, 19652.0, 1.25, 4, 6, 0.5, str(tmp_path), 1000,
                                                              str(tmp_path), None, 'INFO', var_ckpt, var_thr, pause, stop))
        p.start()
        try:
            first = [q.get(timeout=180) for _ in range(10)]
            assert all(st['training_steps'] == 0 for _, st in first)
            var_ckpt.value = str(ckpt).encode('utf-8')
            seen7 = False
            for _ in range(400):
                seq, st = q.get(timeout=180)
                assert seq[0].state.shape == (17, 9, 9) and seq[0].pi_prob.shape == (81,) and seq[0].pi_prob.dtype == np.float32
                assert st['game_result'] in ('B+1.0', 'W+1.0', 'DRAW') and 'num_passes' not in st
                if st['training_steps'] == 7:
                    seen7 = True
                    break
            assert seen7
        finally:
            stop.set()
            for _ in range(200):  # drain so the child's blocking put can finish
                try:
                    q.get(timeout=0.5)
                except Exception:
                    break
            p.join(timeout=120)
            if p.is_alive():
                p.terminate()
    assert os.path.exists(tmp_path / 'actor3.csv')
