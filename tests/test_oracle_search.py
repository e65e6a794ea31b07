"""Oracle search / net / game loop vs traces recorded from the reference (CPU)."""
import os

import numpy as np
import pytest
import torch

from fake_eval import make_fake_eval
from oracle import net as onet
from oracle.boards import GoBoard, GomokuBoard
from oracle.search import search
from oracle.selfplay import play_one_game


def _env(game, max_steps=None):
    return GoBoard(9, 7.5, 8, max_steps) if game == 'go9' else GomokuBoard(13, 5, 8)


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_search_traces(golden_dir, game):
    z = np.load(os.path.join(golden_dir, f'mcts_{game}.npz'))
    A = 82 if game == 'go9' else 169
    ev = make_fake_eval(A)
    for name in z['names']:
        prefix, plies, sims, par, noise, det, warm_steps, reuse, seed = (int(v) for v in z[f'{name}/cfg'])
        np.random.seed(seed)
        env = _env(game)
        for a in z[f'{name}/prefix']:
            env.step(int(a))
        root = None
        for ply in range(len(z[f'{name}/move'])):
            warm = env.steps <= warm_steps
            assert int(warm) == int(z[f'{name}/warm'][ply])
            mv, pi, rq, cq, nxt, child_N = search(env, ev, root if reuse else None, 19652.0, 1.25, sims, par, bool(noise), warm, bool(det))
            tag = f'{game}/{name}/ply{ply}'
            np.testing.assert_array_equal(child_N, z[f'{name}/child_N'][ply], err_msg=tag)
            assert mv == int(z[f'{name}/move'][ply]), tag
            np.testing.assert_array_equal(np.asarray(pi, dtype=np.float64), z[f'{name}/pi'][ply], err_msg=tag)
            assert float(rq) == z[f'{name}/root_q'][ply] and float(cq) == z[f'{name}/child_q'][ply], tag
            assert int(nxt is None) == int(z[f'{name}/next_is_none'][ply]), tag
            _, r, d, _ = env.step(mv)
            assert r == z[f'{name}/reward'][ply] and int(d) == int(z[f'{name}/done'][ply])
            root = nxt


def test_net_forward(golden_dir):
    z = np.load(os.path.join(golden_dir, 'net.npz'))
    torch.set_num_threads(1)
    for tag in ('go9_small', 'gomoku13_small'):
        sd = {k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + '/sd/')}
        gomoku = bool(z[tag + '/cfg'][5])
        logits, v = onet.forward(sd, torch.from_numpy(z[tag + '/x']).float(), gomoku)
        np.testing.assert_allclose(logits.numpy(), z[tag + '/logits'], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(v.numpy(), z[tag + '/v'], rtol=1e-5, atol=1e-6)


def _small_net_sd(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, 'net.npz'))
    return {k[len(tag) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + '/sd/')}


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_game_loop_traces(golden_dir, game):
    """play_and_record_one_game end to end: same seeded numpy stream => same moves, pi, z, stats."""
    z = np.load(os.path.join(golden_dir, f'pipeline_{game}.npz'))
    torch.set_num_threads(1)
    sd = _small_net_sd(golden_dir, f'{game}_small')
    ev = onet.make_eval_func(sd, gomoku=(game != 'go9'))
    for name in sorted({k.split('/')[0] for k in z.files if '/' in k}):
        nb, nf, fc, sims, par, warm, chk, resign_disabled, max_steps, seed = (int(v) for v in z[f'{name}/cfg'])
        thr = float(z[f'{name}/thr'][0])
        import random
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        env = _env(game, max_steps)
        seq, stats = play_one_game(env, ev, sims, par, bool(resign_disabled), 19652.0, 1.25, warm, chk, thr)
        np.testing.assert_array_equal(np.array(env.history, dtype=np.int32), z[f'{name}/history'], err_msg=name)
        assert stats['game_length'] == int(z[f'{name}/game_length'][0])
        assert stats['game_result'] == str(z[f'{name}/result'][0])
        assert repr({k: stats[k] for k in sorted(stats)}) == str(z[f'{name}/stats_repr'][0])
        np.testing.assert_array_equal(np.stack([s for s, _, _ in seq]), z[f'{name}/states'])
        np.testing.assert_allclose(np.stack([p for _, p, _ in seq]), z[f'{name}/pis'], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(np.array([v for _, _, v in seq], dtype=np.float32), z[f'{name}/values'])
