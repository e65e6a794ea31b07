#!/usr/bin/env python
"""Generate the committed golden vectors by RUNNING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py            # regenerates every tests/golden/*.npz

The reference (michaelnny/alpha_zero @ /root/reference) is imported through the stub packages in
tests/shims (gym / sgf / snappy are absent from the image).  `/root/reference` does not exist on the
GPU box, so nothing at test time reads it: the tests consume only the .npz files written here.
Recorded with every file: numpy / torch versions (the reference pins numpy 1.24 / torch 2.0.1;
NEP-50 scalar promotion of numpy>=2 changes `Node.child_U` inputs by <=1 ulp(f32)).

Files (all small, compressed):
  go9_selfplay.npz      every reference self-play SGF (games/selfplay_games/go/9x9): move lists, RE[] strings,
                        per-game SHA1 digest of the env trajectory, full trajectory for the first games.
  gomoku13_selfplay.npz same for games/selfplay_games/gomoku/13x13.
  go19_unit.npz         the move sequences the reference's unit tests pin (unit_tests/envs/go_test.py:80-276)
                        replayed on the reference env, with the outcomes the tests assert.
  gomoku_unit.npz       gomoku_test.py style win-detection cases (7x7 board, num_to_win 3/4/5).
  mcts_<game>.npz       uct_search / parallel_uct_search traces under the deterministic fake evaluator.
  net.npz               AlphaZeroNet forward (small random nets, BN statistics randomised) inputs/outputs.
  pipeline_<game>.npz   play_and_record_one_game end-to-end traces (seeded numpy RNG, small CPU net).
  eval_dataset_go9.npz  replay_sgf (core/eval_dataset.py:80) on 410 recorded games: kept / dropped, per-game digests, MISMATCH_GAMES.
"""
import argparse
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
REF = '/root/reference'
SGF_COLS = 'abcdefghijklmnopqrstuvwxyz'


def _setup_path():
    sys.path.insert(0, TESTS)
    sys.path.insert(0, os.path.join(TESTS, 'shims'))
    sys.path.insert(0, REF)


def parse_sgf(text, n):
    """Moves with the same regex the reference uses (core/eval_dataset.py:124)."""
    res = re.search(r'RE\[([^\]]*)\]', text)
    result = res.group(1) if res else ''
    moves = []
    for m in re.findall(r';[BW]\[[a-z]{0,2}\]', text):
        c = m[3:-1]
        if c == '' or (n <= 19 and c == 'tt'):
            moves.append(n * n)
        else:
            moves.append(SGF_COLS.index(c[1]) * n + SGF_COLS.index(c[0]))
    return moves, result


def versions():
    import numpy as np
    import torch

    return np.array([f'numpy={np.__version__}', f'torch={torch.__version__}', 'reference=michaelnny/alpha_zero@41ec8d65'])


class Trajectory:
    """Canonical byte serialisation of an env trajectory; the oracle and the CUDA tests rebuild the same bytes."""

    def __init__(self):
        self.h = hashlib.sha1()
        self.rows = []

    def add(self, env, reward, done):
        import numpy as np

        legal = np.asarray(env.legal_actions).astype(np.uint8)
        board = np.asarray(env.board).astype(np.int8).ravel()
        tail = np.array([int(reward), int(done), int(env.to_play)], dtype=np.int8)
        self.h.update(legal.tobytes())
        self.h.update(board.tobytes())
        self.h.update(tail.tobytes())
        self.rows.append((legal, board, tail))


def replay_corpus(env, files, n, full_count):
    import numpy as np

    all_moves, offsets, results_sgf, results_env, digests, lens_played = [], [0], [], [], [], []
    full = {'legal': [], 'board': [], 'tail': [], 'obs_last': [], 'game': []}
    for gi, path in enumerate(files):
        moves, result = parse_sgf(open(path).read(), n)
        env.reset()
        tr = Trajectory()
        played = 0
        for a in moves:
            if env.is_game_over():
                break
            _, reward, done, _ = env.step(a)
            tr.add(env, reward, done)
            played += 1
        all_moves.extend(moves)
        offsets.append(len(all_moves))
        results_sgf.append(result)
        results_env.append(env.get_result_string() if env.is_game_over() else '')
        digests.append(tr.h.hexdigest())
        lens_played.append(played)
        if gi < full_count:
            for (lg, bd, tl) in tr.rows:
                full['legal'].append(lg)
                full['board'].append(bd)
                full['tail'].append(tl)
                full['game'].append(gi)
            full['obs_last'].append(env.observation().astype(np.int8))
    out = dict(
        moves=np.array(all_moves, dtype=np.int16),
        offsets=np.array(offsets, dtype=np.int32),
        result_sgf=np.array(results_sgf),
        result_env=np.array(results_env),
        digest=np.array(digests),
        played=np.array(lens_played, dtype=np.int32),
        full_legal=np.stack(full['legal']),
        full_board=np.stack(full['board']),
        full_tail=np.stack(full['tail']),
        full_game=np.array(full['game'], dtype=np.int32),
        full_obs_last=np.stack(full['obs_last']),
        versions=versions(),
    )
    return out


# ---------------------------------------------------------------------------------------------
def part_go9_selfplay():
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    d = os.path.join(REF, 'games/selfplay_games/go/9x9')
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith('.sgf'))
    env = GoEnv(komi=7.5, num_stack=8)
    out = replay_corpus(env, files, 9, full_count=24)
    scored = [(s, e) for s, e in zip(out['result_sgf'], out['result_env']) if not s.endswith('+R')]
    agree = sum(1 for s, e in scored if s == e)
    print(f'go9: {len(files)} games, {len(scored)} scored, {agree} RE[] reproduced')
    np.savez_compressed(os.path.join(HERE, 'go9_selfplay.npz'), **out)


def part_gomoku13_selfplay():
    import numpy as np
    from alpha_zero.envs.gomoku import GomokuEnv

    d = os.path.join(REF, 'games/selfplay_games/gomoku/13x13')
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith('.sgf'))
    env = GomokuEnv(board_size=13, num_stack=8)
    out = replay_corpus(env, files, 13, full_count=24)
    agree = sum(1 for s, e in zip(out['result_sgf'], out['result_env']) if s == e)
    print(f'gomoku13: {len(files)} games, {agree} RE[] reproduced')
    np.savez_compressed(os.path.join(HERE, 'gomoku13_selfplay.npz'), **out)


def part_go19_unit():
    """The sequences of unit_tests/envs/go_test.py, outcomes read off the reference env (19x19)."""
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    cases = {}

    def run(name, gtp_moves, probe=None, max_steps=None, num_stack=8):
        env = GoEnv(num_stack=num_stack) if max_steps is None else GoEnv(max_steps=max_steps, num_stack=num_stack)
        env.reset()
        acts, rewards, dones = [], [], []
        for g in gtp_moves:
            a = env.resign_move if g == 'RESIGN' else env.gtp_to_action(g, check_illegal=False)
            _, r, dn, _ = env.step(a)
            acts.append(a)
            rewards.append(r)
            dones.append(dn)
        cases[name + '/actions'] = np.array(acts, dtype=np.int32)
        cases[name + '/rewards'] = np.array(rewards, dtype=np.float32)
        cases[name + '/dones'] = np.array(dones, dtype=np.uint8)
        cases[name + '/legal'] = np.asarray(env.legal_actions).astype(np.uint8)
        cases[name + '/board'] = np.asarray(env.board).astype(np.int8)
        cases[name + '/obs'] = env.observation().astype(np.int8)
        cases[name + '/winner'] = np.array([0 if env.winner is None else env.winner], dtype=np.int32)
        cases[name + '/result'] = np.array([env.get_result_string() if env.is_game_over() else ''])
        cases[name + '/steps'] = np.array([env.steps], dtype=np.int32)
        if probe is not None:
            cases[name + '/probe'] = np.array([env.gtp_to_action(probe, check_illegal=False)], dtype=np.int32)
            cases[name + '/probe_legal'] = np.array([env.legal_actions[env.gtp_to_action(probe, check_illegal=False)]], dtype=np.uint8)
        return env

    # go_test.py:80-112 suicide
    run('suicide_B1', ('A3', 'A2', 'B2', 'A1', 'C1'), probe='B1')
    run('suicide_F4', ('D3', 'A1', 'D4', 'A2', 'D5', 'A3', 'E3', 'A4', 'E5', 'A5', 'F3', 'A6', 'F5', 'E4', 'G4'), probe='F4')
    # go_test.py:114-127 ko
    ko = []
    for b, w in zip(['A4', 'B4', 'C3', 'C1', 'D2'], ['A2', 'A3', 'B1', 'B3', 'C2']):
        ko += [b, w]
    run('ko_C2', tuple(ko) + ('B2',), probe='C2')
    # go_test.py:129-173 game over by resign / two passes / max steps
    env = GoEnv(num_stack=8)
    gtp = [env.action_to_gtp(i) for i in range(4)]
    run('over_resign', tuple(gtp) + ('RESIGN',))
    run('over_pass', tuple(gtp) + ('PASS', 'PASS'))
    seq = []
    for i in range(4):
        seq += [env.action_to_gtp(i), 'PASS']
    run('pass_steps', tuple(seq))
    for ms in (31, 101):
        run(f'max_steps_{ms}', tuple(env.action_to_gtp(i) for i in range(ms)), max_steps=ms)
    # go_test.py:175-209 score known answers
    run('score_black', ('C1', 'A1', 'B2', 'A2', 'A3', 'PASS', 'PASS'))
    run('score_white', ('A1', 'D2', 'A2', 'C3', 'A3', 'C4', 'B1', 'D5', 'D3', 'E4', 'D4', 'E3', 'PASS', 'PASS'))
    # go_test.py:211-220 winner by resign
    for k in (6, 9):
        run(f'resign_after_{k}', tuple(env.action_to_gtp(i) for i in range(k)) + ('RESIGN',))
    # go_test.py:222-276 stacked observation
    st = []
    for b, w in zip(['B2', 'C3', 'C1', 'B3'], ['A3', 'A1', 'C2', 'B1']):
        st += [b, w]
    run('stacked_obs', tuple(st))
    run('stacked_obs_4', tuple(st), num_stack=4)
    cases['versions'] = versions()
    np.savez_compressed(os.path.join(HERE, 'go19_unit.npz'), **cases)
    print('go19_unit:', len([k for k in cases if k.endswith('/actions')]), 'cases')


def part_gomoku_unit():
    """gomoku_test.py:17-161 style: lines in 4 directions x both colours, num_to_win 3/4/5, on 7x7 (seeded fillers)."""
    import numpy as np
    from alpha_zero.envs.gomoku import GomokuEnv

    rng = np.random.RandomState(7)
    cases = {}
    idx = 0
    for num_to_win in (3, 4, 5):
        for (dr, dc) in ((0, 1), (1, 0), (1, 1), (1, -1)):
            for colour in (0, 1):
                for trial in range(3):
                    n = 7
                    env = GomokuEnv(board_size=n, num_to_win=num_to_win, num_stack=8)
                    env.reset()
                    r0 = rng.randint(0, n - (num_to_win - 1) * abs(dr))
                    c0 = rng.randint(0, n - (num_to_win - 1)) if dc >= 0 else rng.randint(num_to_win - 1, n)
                    line = [(r0 + i * dr) * n + (c0 + i * dc) for i in range(num_to_win)]
                    acts, rewards, dones = [], [], []
                    li = 0
                    first = True
                    while not env.is_game_over():
                        mover_is_line = (env.steps % 2) == colour
                        if mover_is_line and li < len(line):
                            a = line[li]
                            li += 1
                        else:
                            cand = [a for a in np.flatnonzero(env.legal_actions) if a not in line]
                            if not cand:
                                break
                            a = int(rng.choice(cand))
                        _, r, dn, _ = env.step(a)
                        acts.append(a)
                        rewards.append(r)
                        dones.append(dn)
                    name = f'case{idx}'
                    idx += 1
                    cases[name + '/cfg'] = np.array([n, num_to_win], dtype=np.int32)
                    cases[name + '/actions'] = np.array(acts, dtype=np.int32)
                    cases[name + '/rewards'] = np.array(rewards, dtype=np.float32)
                    cases[name + '/dones'] = np.array(dones, dtype=np.uint8)
                    cases[name + '/winner'] = np.array([0 if env.winner is None else env.winner], dtype=np.int32)
                    cases[name + '/result'] = np.array([env.get_result_string()])
                    cases[name + '/obs'] = env.observation().astype(np.int8)
    # a full-board draw on 5x5 with num_to_win=5 is hard to script; use 3x3 num_to_win=3 known draw
    env = GomokuEnv(board_size=3, num_to_win=3, num_stack=8)
    env.reset()
    acts, rewards, dones = [], [], []
    for a in (0, 1, 2, 4, 3, 5, 7, 6, 8):
        _, r, dn, _ = env.step(a)
        acts.append(a)
        rewards.append(r)
        dones.append(dn)
    name = f'case{idx}'
    cases[name + '/cfg'] = np.array([3, 3], dtype=np.int32)
    cases[name + '/actions'] = np.array(acts, dtype=np.int32)
    cases[name + '/rewards'] = np.array(rewards, dtype=np.float32)
    cases[name + '/dones'] = np.array(dones, dtype=np.uint8)
    cases[name + '/winner'] = np.array([0 if env.winner is None else env.winner], dtype=np.int32)
    cases[name + '/result'] = np.array([env.get_result_string()])
    cases[name + '/obs'] = env.observation().astype(np.int8)
    cases['versions'] = versions()
    np.savez_compressed(os.path.join(HERE, 'gomoku_unit.npz'), **cases)
    print('gomoku_unit:', idx + 1, 'cases')


# ---------------------------------------------------------------------------------------------
def _mcts_traces(make_env, num_actions, game_tag):
    """Run the reference searches ply by ply with subtree reuse, exactly like play_and_record_one_game drives them
    (pipeline.py:314-343), under the fake evaluator; capture everything a parity test needs."""
    import numpy as np
    from alpha_zero.core import mcts_v2
    from fake_eval import make_fake_eval

    eval_func = make_fake_eval(num_actions)
    out = {}
    captured = []
    real_dirichlet = np.random.dirichlet

    def spy_dirichlet(alphas):
        x = real_dirichlet(alphas)
        captured.append(np.array(x, dtype=np.float64))
        return x

    # (name, prefix_random_moves, plies, sims, parallel, noise, deterministic, warm_up_steps, reuse, seed)
    scenarios = [
        ('serial_det', 0, 12, 48, 1, False, True, 4, True, 11),
        ('serial_noise', 6, 8, 40, 1, True, False, 4, True, 12),
        ('par_det', 0, 16, 64, 8, False, True, 6, True, 13),
        ('par_noise', 4, 16, 64, 8, True, False, 6, True, 14),
        ('par_noise_late', 46 if game_tag == 'go9' else 60, 14, 96, 8, True, False, 2, True, 15),
        ('par_det_noreuse', 10, 6, 50, 4, False, True, 0, False, 16),
        ('par_full', 2, 3, 400, 8, True, False, 30, True, 17),
    ]
    names = []
    for (name, prefix, plies, sims, par, noise, det, warm_steps, reuse, seed) in scenarios:
        np.random.seed(seed)
        env = make_env()
        env.reset()
        rng = np.random.RandomState(seed + 1000)
        prefix_moves = []
        while len(prefix_moves) < prefix and not env.is_game_over():
            legal = np.flatnonzero(env.legal_actions)
            if env.has_pass_move:  # keep the prefix a real fight, pass only when nothing else is left
                legal = legal[legal != env.pass_move] if len(legal) > 1 else legal
            a = int(rng.choice(legal))
            env.step(a)
            prefix_moves.append(a)
        root = None
        rec = {k: [] for k in ('move', 'pi', 'root_q', 'child_q', 'child_N', 'child_W', 'root_N', 'noise', 'warm', 'reward', 'done', 'next_is_none')}
        c_base, c_init = 19652.0, 1.25
        mcts_v2.np.random.dirichlet = spy_dirichlet
        try:
            for ply in range(plies):
                if env.is_game_over():
                    break
                warm = env.steps <= warm_steps
                del captured[:]
                # keep a handle on the root the search will use, to read its arrays afterwards
                holder = {}
                orig_gsp = mcts_v2.generate_search_policy

                def spy_gsp(child_N, temperature, legal_actions, _h=holder):
                    _h['child_N'] = np.array(child_N, dtype=np.float32)
                    return orig_gsp(child_N, temperature, legal_actions)

                mcts_v2.generate_search_policy = spy_gsp
                try:
                    if par > 1:
                        mv, pi, rq, cq, nxt = mcts_v2.parallel_uct_search(
                            env, eval_func, root if reuse else None, c_base, c_init, sims, par, noise, warm, det
                        )
                    else:
                        mv, pi, rq, cq, nxt = mcts_v2.uct_search(env, eval_func, root if reuse else None, c_base, c_init, sims, noise, warm, det)
                finally:
                    mcts_v2.generate_search_policy = orig_gsp
                rec['move'].append(int(mv))
                rec['pi'].append(np.asarray(pi, dtype=np.float64))
                rec['root_q'].append(float(rq))
                rec['child_q'].append(float(cq))
                rec['child_N'].append(holder['child_N'])
                rec['root_N'].append(float(holder['child_N'].sum()))
                rec['noise'].append(captured[0] if captured else np.zeros(num_actions))
                rec['warm'].append(int(warm))
                rec['next_is_none'].append(int(nxt is None))
                _, r, dn, _ = env.step(int(mv))
                rec['reward'].append(float(r))
                rec['done'].append(int(dn))
                root = nxt
        finally:
            mcts_v2.np.random.dirichlet = real_dirichlet
        names.append(name)
        out[name + '/cfg'] = np.array([prefix, plies, sims, par, int(noise), int(det), warm_steps, int(reuse), seed], dtype=np.int32)
        out[name + '/prefix'] = np.array(prefix_moves, dtype=np.int32)
        out[name + '/move'] = np.array(rec['move'], dtype=np.int32)
        out[name + '/pi'] = np.stack(rec['pi'])
        out[name + '/root_q'] = np.array(rec['root_q'], dtype=np.float64)
        out[name + '/child_q'] = np.array(rec['child_q'], dtype=np.float64)
        out[name + '/child_N'] = np.stack(rec['child_N'])
        out[name + '/noise'] = np.stack(rec['noise'])
        out[name + '/warm'] = np.array(rec['warm'], dtype=np.int32)
        out[name + '/reward'] = np.array(rec['reward'], dtype=np.float32)
        out[name + '/done'] = np.array(rec['done'], dtype=np.int32)
        out[name + '/next_is_none'] = np.array(rec['next_is_none'], dtype=np.int32)
        print(f'  {game_tag}/{name}: {len(rec["move"])} plies, moves {rec["move"][:8]}...')
    out['names'] = np.array(names)
    out['versions'] = versions()
    return out


def part_mcts_go9():
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    out = _mcts_traces(lambda: GoEnv(komi=7.5, num_stack=8), 82, 'go9')
    np.savez_compressed(os.path.join(HERE, 'mcts_go9.npz'), **out)


def part_mcts_gomoku13():
    import numpy as np
    from alpha_zero.envs.gomoku import GomokuEnv

    out = _mcts_traces(lambda: GomokuEnv(board_size=13, num_stack=8), 169, 'gomoku13')
    np.savez_compressed(os.path.join(HERE, 'mcts_gomoku13.npz'), **out)


# ---------------------------------------------------------------------------------------------
def _randomise_bn(net, gen):
    import torch

    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) * 1.5 + 0.25)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=gen) * 1.0 + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=gen) * 0.2)


def part_net():
    import numpy as np
    import torch
    from alpha_zero.core.network import AlphaZeroNet

    torch.set_num_threads(1)
    out = {}
    # (tag, board, actions, blocks, filters, fc, gomoku, batch)
    specs = [
        ('go9_small', 9, 82, 2, 32, 32, False, 6),
        ('go9_c2', 9, 82, 10, 128, 128, False, 3),
        ('gomoku13_small', 13, 169, 2, 32, 48, True, 5),
        ('gomoku13_c4', 13, 169, 6, 64, 64, True, 2),
    ]
    for (tag, n, a, nb, nf, fc, gomoku, batch) in specs:
        torch.manual_seed(123)
        gen = torch.Generator().manual_seed(321)
        net = AlphaZeroNet((17, n, n), a, nb, nf, fc, gomoku)
        _randomise_bn(net, gen)
        net.eval()
        x = (torch.rand((batch, 17, n, n), generator=gen) < 0.3).to(torch.int8)
        x[:, 16] = (torch.arange(batch) % 2).view(-1, 1, 1).to(torch.int8)
        with torch.no_grad():
            logits, v = net(x.float())
            pi = torch.softmax(logits, dim=-1)
        out[tag + '/cfg'] = np.array([n, a, nb, nf, fc, int(gomoku)], dtype=np.int32)
        out[tag + '/x'] = x.numpy()
        out[tag + '/logits'] = logits.numpy()
        out[tag + '/pi'] = pi.numpy()
        out[tag + '/v'] = v.numpy()
        if nf <= 32:  # the small nets travel with their weights; the big ones are re-created from the seed by my own module
            for k, t in net.state_dict().items():
                out[tag + '/sd/' + k] = t.numpy()
        else:
            sd = net.state_dict()
            out[tag + '/sd_keys'] = np.array(list(sd.keys()))
            out[tag + '/sd_sum'] = np.array([float(t.double().sum()) for t in sd.values()])
    out['versions'] = versions()
    np.savez_compressed(os.path.join(HERE, 'net.npz'), **out)
    print('net: ok')


def _pipeline_trace(make_env, n, a, gomoku, tag):
    import logging

    import numpy as np
    import torch
    from alpha_zero.core.network import AlphaZeroNet
    from alpha_zero.core.pipeline import create_mcts_player, play_and_record_one_game, set_seed

    torch.set_num_threads(1)
    out = {}
    logger = logging.getLogger('golden')
    # (name, blocks, filters, fc, sims, parallel, warm_up_steps, check_resign_after, resign_disabled, resign_thr, max_steps, seed)
    runs = [
        ('par', 2, 32, 32 if not gomoku else 48, 24, 4, 6, 4, True, -1.0, 40, 5),
        ('serial', 2, 32, 32 if not gomoku else 48, 16, 1, 4, 4, True, -1.0, 30, 6),
    ]
    if not gomoku:
        runs.append(('par_resign', 2, 32, 32, 24, 4, 4, 2, False, 0.95, 40, 7))
    for (name, nb, nf, fc, sims, par, warm, chk, resign_disabled, thr, max_steps, seed) in runs:
        torch.manual_seed(123)
        gen = torch.Generator().manual_seed(321)
        net = AlphaZeroNet((17, n, n), a, nb, nf, fc, gomoku)
        _randomise_bn(net, gen)
        net.eval()
        env = make_env(max_steps)
        set_seed(seed)
        player = create_mcts_player(net, torch.device('cpu'), sims, par, root_noise=True, deterministic=False)
        seq, stats = play_and_record_one_game(env, player, resign_disabled, 19652.0, 1.25, warm, chk, thr, logger)
        out[name + '/cfg'] = np.array([nb, nf, fc, sims, par, warm, chk, int(resign_disabled), max_steps, seed], dtype=np.int32)
        out[name + '/thr'] = np.array([thr], dtype=np.float64)
        out[name + '/states'] = np.stack([t.state for t in seq]).astype(np.int8)
        out[name + '/pis'] = np.stack([np.asarray(t.pi_prob, dtype=np.float64) for t in seq])
        out[name + '/values'] = np.array([t.value for t in seq], dtype=np.float32)
        hist = [m.move for m in env.history]
        out[name + '/history'] = np.array(hist, dtype=np.int32)
        out[name + '/last_move'] = np.array([env.last_move], dtype=np.int32)
        out[name + '/result'] = np.array([stats['game_result']])
        out[name + '/game_length'] = np.array([stats['game_length']], dtype=np.int32)
        out[name + '/stats_keys'] = np.array(sorted(stats.keys()))
        out[name + '/stats_repr'] = np.array([repr({k: stats[k] for k in sorted(stats)})])
        print(f'  {tag}/{name}: {stats}')
    out['versions'] = versions()
    return out


def part_pipeline_go9():
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    out = _pipeline_trace(lambda ms: GoEnv(komi=7.5, num_stack=8, max_steps=ms), 9, 82, False, 'go9')
    np.savez_compressed(os.path.join(HERE, 'pipeline_go9.npz'), **out)


def part_pipeline_gomoku13():
    import numpy as np
    from alpha_zero.envs.gomoku import GomokuEnv

    out = _pipeline_trace(lambda ms: GomokuEnv(board_size=13, num_stack=8), 13, 169, True, 'gomoku13')
    np.savez_compressed(os.path.join(HERE, 'pipeline_gomoku13.npz'), **out)



def _random_games(env, n_games, plies, seed, n):
    """Seeded random legal play (passes allowed only when nothing else is legal) through the reference env."""
    import numpy as np

    rng = np.random.RandomState(seed)
    all_moves, offsets, digests, results = [], [0], [], []
    for _ in range(n_games):
        env.reset()
        tr = Trajectory()
        for _ply in range(plies):
            if env.is_game_over():
                break
            legal = np.flatnonzero(env.legal_actions)
            if env.has_pass_move and len(legal) > 1:
                legal = legal[legal != env.pass_move]
            a = int(rng.choice(legal))
            _, r, d, _ = env.step(a)
            tr.add(env, r, d)
            all_moves.append(a)
        offsets.append(len(all_moves))
        digests.append(tr.h.hexdigest())
        results.append(env.get_result_string() if env.is_game_over() else '')
    return dict(moves=np.array(all_moves, dtype=np.int16), offsets=np.array(offsets, dtype=np.int32), digest=np.array(digests),
                result_env=np.array(results), versions=versions())


def part_random_go19():
    """Big-board coverage: 12 random 19x19 games of up to 420 plies (captures, ko, suicide masks on 361 cells)."""
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    out = _random_games(GoEnv(komi=7.5, num_stack=8), 12, 420, 19, 19)
    np.savez_compressed(os.path.join(HERE, 'random_go19.npz'), **out)
    print('random_go19: ok')


def part_random_go13():
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    out = _random_games(GoEnv(komi=5.5, num_stack=8, max_steps=300), 24, 300, 13, 13)
    np.savez_compressed(os.path.join(HERE, 'random_go13.npz'), **out)
    print('random_go13: ok')


def part_random_go9_full():
    """Random 9x9 games played to the end (max_steps / double pass): scoring of crowded final positions."""
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    out = _random_games(GoEnv(komi=7.5, num_stack=8), 60, 400, 9, 9)
    np.savez_compressed(os.path.join(HERE, 'random_go9.npz'), **out)
    print('random_go9: ok', sum(1 for r in out['result_env'] if r))


def part_random_gomoku15():
    import numpy as np
    from alpha_zero.envs.gomoku import GomokuEnv

    out = _random_games(GomokuEnv(board_size=15, num_stack=8), 40, 225, 15, 15)
    np.savez_compressed(os.path.join(HERE, 'random_gomoku15.npz'), **out)
    print('random_gomoku15: ok')


def part_pro_go9():
    """Human games (games/pro_games/go/9x9): tougher fights than self-play.  First 1500 files, digests only."""
    import numpy as np
    from alpha_zero.envs.go import GoEnv

    d = os.path.join(REF, 'games/pro_games/go/9x9')
    files = sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith('.sgf'))[:1500]
    env = GoEnv(komi=7.5, num_stack=8)
    all_moves, offsets, digests, played = [], [0], [], []
    for path in files:
        try:
            moves, _ = parse_sgf(open(path, errors='ignore').read(), 9)
        except Exception:
            moves = []
        env.reset()
        tr = Trajectory()
        ok = []
        for a in moves:
            if env.is_game_over() or a > 81 or env.legal_actions[a] != 1:
                break
            _, r, dn, _ = env.step(a)
            tr.add(env, r, dn)
            ok.append(a)
        all_moves.extend(ok)
        offsets.append(len(all_moves))
        digests.append(tr.h.hexdigest())
        played.append(len(ok))
    out = dict(moves=np.array(all_moves, dtype=np.int16), offsets=np.array(offsets, dtype=np.int32), digest=np.array(digests),
               played=np.array(played, dtype=np.int32), versions=versions())
    np.savez_compressed(os.path.join(HERE, 'pro_go9.npz'), **out)
    print('pro_go9:', len(files), 'games,', len(all_moves), 'moves')



def part_eval_dataset_go9():
    """core/eval_dataset.py:80 replay_sgf run by the reference itself (through the `sgf` stand-in of tests/shims) on a mixed
    list of files: human 9x9 games, the CrazyStone matches (scored results: MISMATCH_GAMES statistics) and self-play records
    (same two player names: duplicate and games-per-player filters).  The SGF texts travel inside the fixture."""
    import logging
    import numpy as np
    from alpha_zero.core import eval_dataset as ed

    def pick(sub, k):
        d = os.path.join(REF, 'games', sub)
        return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith('.sgf')][:k]

    files = pick('pro_games/go/9x9', 260) + pick('9x9_matches/crazystone_vs_az', 20) + pick('selfplay_games/go/9x9', 130)
    logger = logging.getLogger('golden')
    logger.setLevel(logging.CRITICAL)
    texts, valid, counts, digests = [], [], [], []
    for path in files:
        texts.append(open(path).read())
        h = ed.replay_sgf(path, 8, logger)
        valid.append(h is not None)
        sha = hashlib.sha1()
        for obs, pi, v in (h or []):
            assert obs.dtype == np.int8 and obs.shape == (17, 9, 9)
            sha.update(obs.tobytes())
            sha.update(np.int32(int(np.argmax(pi))).tobytes())
            sha.update(np.float32(v).tobytes())
        counts.append(len(h) if h is not None else -1)
        digests.append(sha.hexdigest())
    mm = ed.MISMATCH_GAMES
    out = dict(names=np.array([os.path.relpath(f, os.path.join(REF, 'games')) for f in files]), texts=np.array(texts), valid=np.array(valid),
               counts=np.array(counts, dtype=np.int32), digest=np.array(digests), mismatch_keys=np.array(list(mm.keys())),
               mismatch_values=np.array(list(mm.values()), dtype=np.int32), versions=versions())
    np.savez_compressed(os.path.join(HERE, 'eval_dataset_go9.npz'), **out)
    print('eval_dataset_go9:', len(files), 'files,', int(np.sum(valid)), 'valid,', int(np.sum(np.maximum(counts, 0))), 'positions,', dict(mm))


def part_transform():
    """utils/transformation.py:160 (learner-side augmentation): every supported transformation applied by
    the reference itself to random (state, pi) batches, with and without the pass column."""
    import numpy as np
    import torch
    from alpha_zero.utils import transformation as tr

    out = {}
    gen = torch.Generator().manual_seed(5)
    for tag, n, has_pass in (('go9', 9, True), ('gomoku13', 13, False), ('go19', 19, True)):
        a = n * n + (1 if has_pass else 0)
        st = (torch.rand((6, 17, n, n), generator=gen) < 0.3).to(torch.float32)
        pi = torch.rand((6, a), generator=gen)
        pi = pi / pi.sum(dim=1, keepdim=True)
        v = torch.rand((6,), generator=gen) * 2 - 1
        out[tag + '/state'] = st.numpy().astype(np.int8)
        out[tag + '/pi'] = pi.numpy()
        out[tag + '/value'] = v.numpy()
        for name in tr.TRANSFORMATIONS:
            s2, p2, v2 = tr.SUPPORTED_TRANSFORMATIONS[name](st, pi, v)
            out[f'{tag}/{name}/state'] = s2.numpy().astype(np.int8)
            out[f'{tag}/{name}/pi'] = p2.numpy()
            assert torch.equal(v2, v)
    out['names'] = np.array(tr.TRANSFORMATIONS)
    out['versions'] = versions()
    np.savez_compressed(os.path.join(HERE, 'transform.npz'), **out)
    print('transform: ok', tr.TRANSFORMATIONS)


def _corpus_positions(env, files, n, per_game, limit, seed):
    """Observations at random plies of recorded games replayed on the reference env (int8 [k, 17, n, n]) + the ply numbers."""
    import numpy as np

    rng = np.random.RandomState(seed)
    obs, plies = [], []
    for path in files:
        try:
            moves, _ = parse_sgf(open(path, errors='ignore').read(), n)
        except Exception:
            continue
        if len(moves) < 4:
            continue
        want = set(rng.choice(len(moves), size=min(per_game, len(moves)), replace=False).tolist())
        env.reset()
        for t, a in enumerate(moves):
            if env.is_game_over() or a >= env.action_dim or env.legal_actions[a] != 1:
                break
            if t in want:
                obs.append(env.observation().astype(np.int8))
                plies.append(t)
            env.step(a)
        if len(obs) >= limit:
            break
    return np.stack(obs[:limit]), np.array(plies[:limit], dtype=np.int32)


def part_net_ckpt():
    """AlphaZeroNet.forward (core/network.py:160) of the reference's TRAINED checkpoints on positions of recorded games:
    checkpoints/go/9x9/training_steps_154000.ckpt (README.md:109, the net SURVEY.md 8d quotes the metric on; 10 blocks x 128
    filters) over 384 positions of self-play + human 9x9 games, and checkpoints/gomoku/13x13/training_steps_219000.ckpt
    (10 blocks x 40 filters, fc 80: training_gomoku.py defaults) over 192 positions of 13x13 self-play games.  The fp32 weights
    travel beside the vectors (ckpt_*.npz: network tensors only, no optimiser state) because the GPU box has no /root/reference;
    the bench's checkpoint-shaped workload (resign rule -0.88) loads the same file."""
    import numpy as np
    import torch
    from alpha_zero.core.network import AlphaZeroNet
    from alpha_zero.envs.go import GoEnv
    from alpha_zero.envs.gomoku import GomokuEnv

    torch.set_num_threads(4)
    out = {}

    def ls(sub):
        d = os.path.join(REF, 'games', sub)
        return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith('.sgf')]

    specs = [
        ('go9_154000', 'checkpoints/go/9x9/training_steps_154000.ckpt', 9, 82, 10, 128, 128, False,
         lambda: GoEnv(komi=7.5, num_stack=8), [(ls('selfplay_games/go/9x9')[::9], 2, 256), (ls('pro_games/go/9x9')[::40], 2, 128)]),
        ('gomoku13_219000', 'checkpoints/gomoku/13x13/training_steps_219000.ckpt', 13, 169, 10, 40, 80, True,
         lambda: GomokuEnv(board_size=13, num_stack=8), [(ls('selfplay_games/gomoku/13x13')[::12], 1, 192)]),
    ]
    for (tag, ck, n, a, nb, nf, fc, gomoku, make_env, sources) in specs:
        sd = torch.load(os.path.join(REF, ck), map_location='cpu')['network']
        net = AlphaZeroNet((17, n, n), a, nb, nf, fc, gomoku)
        net.load_state_dict(sd)
        net.eval()
        parts, plies = [], []
        for si, (files, per_game, limit) in enumerate(sources):
            o, p = _corpus_positions(make_env(), files, n, per_game, limit, seed=7 + si)
            parts.append(o)
            plies.append(p)
        x = np.concatenate(parts)
        with torch.no_grad():
            logits, v = net(torch.from_numpy(x).float())
            pi = torch.softmax(logits, dim=-1)
        out[tag + '/cfg'] = np.array([n, a, nb, nf, fc, int(gomoku)], dtype=np.int32)
        out[tag + '/x'] = x
        out[tag + '/ply'] = np.concatenate(plies)
        out[tag + '/logits'] = logits.numpy()
        out[tag + '/pi'] = pi.numpy()
        out[tag + '/v'] = v.numpy()
        w = {k: t.numpy() for k, t in net.state_dict().items()}
        w['versions'] = versions()
        np.savez_compressed(os.path.join(HERE, f'ckpt_{tag}.npz'), **w)
        print(f'net_ckpt {tag}: {x.shape[0]} positions, mean max-prob {float(pi.max(dim=1).values.mean()):.3f}, mean |v| {float(v.abs().mean()):.3f}')
    out['versions'] = versions()
    np.savez_compressed(os.path.join(HERE, 'net_ckpt.npz'), **out)


PARTS = {
    'net_ckpt': (part_net_ckpt, 9),
    'go9_selfplay': (part_go9_selfplay, 9),
    'gomoku13_selfplay': (part_gomoku13_selfplay, 9),
    'go19_unit': (part_go19_unit, 19),
    'gomoku_unit': (part_gomoku_unit, 9),
    'mcts_go9': (part_mcts_go9, 9),
    'mcts_gomoku13': (part_mcts_gomoku13, 9),
    'net': (part_net, 9),
    'pipeline_go9': (part_pipeline_go9, 9),
    'pipeline_gomoku13': (part_pipeline_gomoku13, 9),
    'random_go19': (part_random_go19, 19),
    'random_go13': (part_random_go13, 13),
    'random_go9': (part_random_go9_full, 9),
    'random_gomoku15': (part_random_gomoku15, 9),
    'pro_go9': (part_pro_go9, 9),
    'transform': (part_transform, 9),
    'eval_dataset_go9': (part_eval_dataset_go9, 9),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--part', default=None)
    ap.add_argument('--only', default=None, help='comma separated subset of parts')
    args = ap.parse_args()
    if args.part:
        _setup_path()
        PARTS[args.part][0]()
        return
    todo = args.only.split(',') if args.only else list(PARTS)
    for name in todo:
        env = dict(os.environ, BOARD_SIZE=str(PARTS[name][1]), OMP_NUM_THREADS='1', MKL_NUM_THREADS='1')
        print(f'[make_golden] {name}')
        subprocess.run([sys.executable, os.path.abspath(__file__), '--part', name], check=True, env=env)


if __name__ == '__main__':
    main()
