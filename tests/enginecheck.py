"""Parity checks of the engine (through the C ABI) against the golden vectors and the oracle.

The same functions run twice: on the host-emulation build of the device code (`-m "not gpu"`, tests/emu)
and on the real CUDA library on a B200 (`-m gpu`).
"""
import os

import numpy as np

from alpha_zero_b200.engine import Engine
from fake_eval import make_fake_eval
from trajectory import Trajectory, parse_corpus

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def ext_to_play(game, scal):
    return scal['to_play']


def replay_corpus(binding, game, stride, batch=64):
    """Replay reference self-play SGFs `batch` games at a time through az_env_step and compare the
    per-game trajectory digests (legal mask, board, reward, done, to_play after every ply)."""
    z = np.load(os.path.join(GOLDEN, f'{game}_selfplay.npz'))
    games = parse_corpus(z)
    idx = list(range(0, len(games), stride))
    kind, n = ('go', 9) if game == 'go9' else ('gomoku', 13)
    eng = Engine(kind, n, num_games=batch, max_simulations=8, max_parallel=1, binding=binding)
    bad = []
    for b0 in range(0, len(idx), batch):
        chunk = idx[b0:b0 + batch]
        slots = np.arange(len(chunk), dtype=np.int32)
        eng.env_reset(slots)
        trs = [Trajectory() for _ in chunk]
        done = [False] * len(chunk)
        played = [0] * len(chunk)
        maxlen = max(len(games[g]) for g in chunk)
        for ply in range(maxlen):
            live = [i for i, g in enumerate(chunk) if ply < len(games[g]) and not done[i]]
            if not live:
                break
            acts = [int(games[chunk[i]][ply]) for i in live]
            r, d = eng.env_step(live, acts)
            for k, i in enumerate(live):
                sc = eng.env_scalars(i)
                trs[i].add(eng.env_legal(i), eng.env_board(i), r[k], d[k], sc['to_play'])
                played[i] += 1
                done[i] = bool(d[k])
        for i, g in enumerate(chunk):
            if trs[i].hexdigest() != str(z['digest'][g]) or played[i] != int(z['played'][g]):
                bad.append(g)
    eng.close()
    return bad, len(idx)


def final_observations(binding, game, count=24):
    z = np.load(os.path.join(GOLDEN, f'{game}_selfplay.npz'))
    games = parse_corpus(z)
    kind, n = ('go', 9) if game == 'go9' else ('gomoku', 13)
    eng = Engine(kind, n, num_games=2, max_simulations=8, max_parallel=1, binding=binding)
    for gi in range(count):
        eng.env_reset([0])
        for a in games[gi]:
            _, d = eng.env_step([0], [int(a)])
            if d[0]:
                break
        np.testing.assert_array_equal(eng.env_observation(0), z['full_obs_last'][gi])
    # deepcopy + pickle round trip of a position
    eng.env_copy(0, 1)
    np.testing.assert_array_equal(eng.env_observation(1), eng.env_observation(0))
    blob = eng.env_export(0)
    eng.env_reset([1])
    eng.env_import(1, blob)
    np.testing.assert_array_equal(eng.env_observation(1), eng.env_observation(0))
    np.testing.assert_array_equal(eng.env_legal(1), eng.env_legal(0))
    eng.close()


def go19_unit(binding):
    z = np.load(os.path.join(GOLDEN, 'go19_unit.npz'))
    names = sorted({k.split('/')[0] for k in z.files if '/' in k})
    for name in names:
        kw = {}
        if name.startswith('max_steps_'):
            kw['max_steps'] = int(name.split('_')[-1])
        if name == 'stacked_obs_4':
            kw['num_stack'] = 4
        eng = Engine('go', 19, num_games=1, max_simulations=8, max_parallel=1, **kw, binding=binding)
        for a, r, d in zip(z[name + '/actions'], z[name + '/rewards'], z[name + '/dones']):
            rr, dd = eng.env_step([0], [int(a)])
            assert rr[0] == r and bool(dd[0]) == bool(d), name
        np.testing.assert_array_equal(eng.env_legal(0), z[name + '/legal'], err_msg=name)
        np.testing.assert_array_equal(eng.env_board(0), z[name + '/board'], err_msg=name)
        np.testing.assert_array_equal(eng.env_observation(0), z[name + '/obs'], err_msg=name)
        sc = eng.env_scalars(0)
        assert sc['winner'] == int(z[name + '/winner'][0]), name
        assert sc['steps'] == int(z[name + '/steps'][0]), name
        res = str(z[name + '/result'][0])
        if sc['done'] and not sc['by_resign']:
            s = eng.env_score(0)
            mine = ('B+%.1f' % s) if s > 0 else (('W+%.1f' % abs(s)) if s < 0 else 'DRAW')
            assert mine == res, (name, mine, res)
        if name + '/probe' in z.files:
            try:
                eng.env_step([0], [int(z[name + '/probe'][0])])
                raise AssertionError(name + ': illegal move accepted')
            except ValueError as ex:
                assert 'Illegal action' in str(ex)
        eng.close()
    eng = Engine('go', 19, num_games=1, max_simulations=8, max_parallel=1, binding=binding)
    for bad in (500, 19 * 19 + 2, 999):
        try:
            eng.env_step([0], [bad])
            raise AssertionError('out-of-range action accepted')
        except ValueError as ex:
            assert 'Invalid action' in str(ex)
    for a in z['over_pass/actions']:
        eng.env_step([0], [int(a)])
    try:
        eng.env_step([0], [6])
        raise AssertionError('step after game over accepted')
    except RuntimeError as ex:
        assert 'Game is over' in str(ex)
    eng.close()


def gomoku_unit(binding):
    z = np.load(os.path.join(GOLDEN, 'gomoku_unit.npz'))
    for name in sorted({k.split('/')[0] for k in z.files if '/' in k}):
        n, k = (int(v) for v in z[name + '/cfg'])
        eng = Engine('gomoku', n, num_games=1, max_simulations=8, max_parallel=1, num_to_win=k, binding=binding)
        for a, r, d in zip(z[name + '/actions'], z[name + '/rewards'], z[name + '/dones']):
            rr, dd = eng.env_step([0], [int(a)])
            assert rr[0] == r and bool(dd[0]) == bool(d), name
        assert eng.env_scalars(0)['winner'] == int(z[name + '/winner'][0])
        np.testing.assert_array_equal(eng.env_observation(0), z[name + '/obs'])
        eng.close()


def run_search(eng, slot, eval_func, reuse, c_base, c_init, sims, par, noise_flag, warm, det, noise=None):
    """Drive one search on one slot through the split-phase ABI with a Python evaluator."""
    eng.search_begin([slot], [1 if reuse else 0], c_base, c_init, sims, par, noise_flag, warm, det, noise=noise)
    while True:
        obs, counts, active = eng.search_select()
        if active == 0:
            break
        if len(obs):
            pri, val = eval_func(obs, True)
            eng.search_apply(np.stack(pri), np.array(val, dtype=np.float32))
        else:
            eng.search_apply(None, None)
    return eng.search_result(slot)


def mcts_traces(binding, game):
    """child_N of every ply of every reference trace must be reproduced EXACTLY; root_Q / best_child_Q too."""
    z = np.load(os.path.join(GOLDEN, f'mcts_{game}.npz'))
    kind, n, A = ('go', 9, 82) if game == 'go9' else ('gomoku', 13, 169)
    ev = make_fake_eval(A)
    eng = Engine(kind, n, num_games=2, max_simulations=416, max_parallel=8, binding=binding)
    checked = 0
    for name in z['names']:
        prefix, plies, sims, par, noise, det, warm_steps, reuse, seed = (int(v) for v in z[f'{name}/cfg'])
        eng.env_reset([0])
        for a in z[f'{name}/prefix']:
            eng.env_step([0], [int(a)])
        has_tree = False
        for ply in range(len(z[f'{name}/move'])):
            warm = bool(z[f'{name}/warm'][ply])
            nz = z[f'{name}/noise'][ply] if noise else None
            res = run_search(eng, 0, ev, bool(reuse) and has_tree, 19652.0, 1.25, sims, par, bool(noise), warm, bool(det), noise=nz)
            tag = f'{game}/{name}/ply{ply}'
            np.testing.assert_array_equal(res['child_N'], z[f'{name}/child_N'][ply], err_msg=tag)
            assert res['root_q'] == z[f'{name}/root_q'][ply], (tag, res['root_q'], z[f'{name}/root_q'][ply])
            np.testing.assert_allclose(res['pi'], z[f'{name}/pi'][ply], rtol=1e-6, atol=1e-9, err_msg=tag)
            mv = int(z[f'{name}/move'][ply])
            if det:
                assert res['argmax'] == mv, tag
            bq, has_tree = eng.search_commit(0, mv)
            assert bq == z[f'{name}/child_q'][ply], (tag, bq, z[f'{name}/child_q'][ply])
            assert int(not has_tree) == int(z[f'{name}/next_is_none'][ply]), tag
            r, d = eng.env_step([0], [mv])
            assert r[0] == z[f'{name}/reward'][ply] and int(d[0]) == int(z[f'{name}/done'][ply]), tag
            checked += 1
    c = eng.counters()
    assert c['errors'] == 0
    eng.close()
    return checked


EXTRA = {  # name -> (game kind, board size, engine kwargs)
    'random_go19': ('go', 19, {}),
    'random_go13': ('go', 13, {'komi': 5.5, 'max_steps': 300}),
    'random_go9': ('go', 9, {}),
    'random_gomoku15': ('gomoku', 15, {}),
    'pro_go9': ('go', 9, {}),
}


def replay_extra(binding, name, stride=1):
    """Random-play / human-game corpora recorded from the reference: all games of a chunk advance in lock step."""
    z = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    games = parse_corpus(z)
    idx = list(range(0, len(games), stride))
    kind, n, kw = EXTRA[name]
    batch = min(64, len(idx))
    eng = Engine(kind, n, num_games=batch, max_simulations=8, max_parallel=1, binding=binding, **kw)
    bad = []
    for b0 in range(0, len(idx), batch):
        chunk = idx[b0:b0 + batch]
        eng.env_reset(np.arange(len(chunk), dtype=np.int32))
        trs = [Trajectory() for _ in chunk]
        done = [False] * len(chunk)
        for ply in range(max(len(games[g]) for g in chunk)):
            live = [i for i, g in enumerate(chunk) if ply < len(games[g]) and not done[i]]
            if not live:
                break
            r, d = eng.env_step(live, [int(games[chunk[i]][ply]) for i in live])
            for k, i in enumerate(live):
                trs[i].add(eng.env_legal(i), eng.env_board(i), r[k], d[k], eng.env_scalars(i)['to_play'])
                done[i] = bool(d[k])
        bad += [g for i, g in enumerate(chunk) if trs[i].hexdigest() != str(z['digest'][g])]
    eng.close()
    return bad, len(idx)


def concurrent_searches(binding, game):
    """Several slots searched in ONE batch (ragged: different positions, different carried subtrees, different leaf counts)
    must each behave exactly as if searched alone: slot 0 replays a reference trace from its first ply, slots 1 and 3 start
    from later positions of the same game and are compared with the oracle run from the same start.  Checks slot
    independence, subtree reuse per slot and the leaf-row compaction order."""
    from oracle.boards import GoBoard, GomokuBoard
    from oracle.search import search

    z = np.load(os.path.join(GOLDEN, f'mcts_{game}.npz'))
    kind, n, A = ('go', 9, 82) if game == 'go9' else ('gomoku', 13, 169)
    ev = make_fake_eval(A)
    nm = 'par_det'
    prefix, plies, sims, par, noise, det, warm_steps, reuse, seed = (int(v) for v in z[f'{nm}/cfg'])
    assert not noise and det and reuse
    eng = Engine(kind, n, num_games=4, max_simulations=sims + 8, max_parallel=par, binding=binding)
    slots, starts = [0, 1, 3], [0, 2, 5]
    moves = [int(m) for m in z[f'{nm}/move']]
    oracles, roots = {}, {}
    for s_, st in zip(slots, starts):
        env = GoBoard(9) if game == 'go9' else GomokuBoard(13)
        for a in [int(a) for a in z[f'{nm}/prefix']] + moves[:st]:
            eng.env_step([s_], [a])
            env.step(a)
        oracles[s_], roots[s_] = env, None
    for step in range(5):
        warm = False  # deterministic play: warm-up only changes the exponent of pi, not the visit counts compared here
        eng.search_begin(slots, [int(step > 0)] * 3, 19652.0, 1.25, sims, par, False, warm, True)
        while True:
            obs, counts, active = eng.search_select()
            if active == 0:
                break
            if len(obs):
                assert counts.sum() == len(obs)
                pri, val = ev(obs, True)
                eng.search_apply(np.stack(pri), np.array(val, dtype=np.float32))
            else:
                eng.search_apply(None, None)
        for s_, st in zip(slots, starts):
            res = eng.search_result(s_)
            mv_o, pi, rq, cq, roots[s_], child_N = search(oracles[s_], ev, roots[s_], 19652.0, 1.25, sims, par, False, warm, True)
            np.testing.assert_array_equal(res['child_N'], child_N, err_msg=f'slot {s_} step {step}')
            assert res['argmax'] == mv_o and res['root_q'] == float(rq)
            if st == 0:
                np.testing.assert_array_equal(res['child_N'], z[f'{nm}/child_N'][step])
                assert mv_o == moves[step]
            bq, has = eng.search_commit(s_, mv_o)
            assert bq == float(cq) and has == (roots[s_] is not None)
            eng.env_step([s_], [mv_o])
            oracles[s_].step(mv_o)
    assert eng.counters()['errors'] == 0
    eng.close()
