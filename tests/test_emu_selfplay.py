"""The device-resident self-play loop (csrc/az_tree.cuh: game_advance / game_new, rings, az_drain_games) on the host-emulation
build with a hash evaluator standing in for the network: every finished game is replayed through the ORACLE board engine and
must be a legal, correctly scored, correctly labelled game."""
import ctypes
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402
from alpha_zero_b200.engine import Engine  # noqa: E402
from oracle.boards import GoBoard, GomokuBoard  # noqa: E402


@pytest.fixture(scope='module')
def emu():
    return Binding(ctypes.CDLL(build_emu.build()))


def _dummy_weights(blocks, filters, fc, planes, A, hw):
    shapes = [(filters * planes * 9,)] + [(filters,)] * 4
    for _ in range(blocks * 2):
        shapes += [(filters * filters * 9,)] + [(filters,)] * 4
    shapes += [(2 * filters,)] + [(2,)] * 4 + [(A * 2 * hw,), (A,)] + [(filters,)] + [(1,)] * 4 + [(fc * hw,), (fc,), (fc,), (1,)]
    return {f't{i}': np.zeros(s, dtype=np.float32) for i, s in enumerate(shapes)}


def _check_game(game, rec, states, pis, zs, moves, max_steps):
    env = GoBoard(9, 7.5, 8, max_steps) if game == 'go' else GomokuBoard(9, 5, 8)
    ln = rec['game_length']
    s0 = rec['first_sample']
    black, white = env.black_player, env.white_player
    obs = env.reset()
    to_plays = []
    reward, done = 0.0, False
    for i in range(ln):
        assert not done
        np.testing.assert_array_equal(states[s0 + i], obs)
        pi = pis[s0 + i]
        assert abs(float(pi.sum()) - 1.0) < 1e-5
        assert np.all(np.asarray(env.legal_actions)[pi > 0] == 1)  # the search policy lives on legal moves only
        mv = int(moves[s0 + i])
        to_plays.append(env.to_play)
        if mv >= 0:
            assert env.legal_actions[mv] == 1
        obs, reward, done, _ = env.step(mv)
    assert done
    assert bool(rec['by_resign']) == (int(moves[s0 + ln - 1]) == -1)
    assert (0 if env.winner is None else env.winner) == rec['winner']
    if game == 'go':
        assert rec['num_passes'] == sum(1 for i in range(ln) if int(moves[s0 + i]) == 81)
        if not rec['by_resign']:
            assert abs(env.score() - rec['score']) < 1e-6
    z = zs[s0:s0 + ln]
    if reward == 0.0:
        assert np.all(z == 0)
    else:  # pipeline.py:349-354
        want = np.array([reward if p == env.last_player else -reward for p in to_plays], dtype=np.float32)
        np.testing.assert_array_equal(z, want)
    if rec['is_marked_for_resign']:
        assert rec['is_resign_disabled'] and rec['marked_resign_player'] in (black, white)
        assert bool(rec['is_could_won']) == (rec['winner'] == rec['marked_resign_player'])
    if rec['by_resign']:
        assert not rec['is_resign_disabled']
    return ln


@pytest.mark.parametrize('game', ['go', 'gomoku'])
def test_selfplay_loop_produces_valid_games(emu, game):
    A = 82 if game == 'go' else 81
    max_steps = 36
    eng = Engine(game, 9, num_games=12, max_simulations=16, max_parallel=4, net=(1, 16, 16), precision='fp32',
                 max_steps=max_steps if game == 'go' else 0, seed=3, sample_ring=700, binding=emu)
    eng.set_weights(_dummy_weights(1, 16, 16, 17, A, 81 if game == 'go' else 169))
    eng.selfplay_begin(12, 4, warm_up_steps=6, check_resign_after_steps=8, resign_threshold=-0.55, disable_resign_ratio=0.5)
    n_games, n_resign, n_marked, total = 0, 0, 0, 0
    for rnd in range(60):
        eng.selfplay_tick(8)
        if rnd == 20:  # knobs can change while games are in flight (var_resign_threshold, pipeline.py:241-246)
            eng.selfplay_update(6, 8, -0.5, 0.5)
        games, states, pis, zs = eng.drain_games()
        mv = eng.last_moves
        for rec in games:
            total += _check_game(game, rec, states, pis, zs, mv, max_steps)
            n_games += 1
            n_resign += rec['by_resign']
            n_marked += rec['is_marked_for_resign']
    c = eng.counters()
    assert c['errors'] == 0 and c['ring_dropped'] == 0 and c['games'] == n_games and c['samples'] == total
    assert n_games >= 12
    if game == 'go':
        assert n_resign > 0 and n_marked > 0  # both branches of the resignation logic were taken
    else:
        assert n_resign == 0
    eng.close()


def test_restart_and_search_update_keep_the_loop_consistent(emu):
    """az_selfplay_restart abandons games without emitting anything; az_selfplay_update with search parameters changes the
    simulations per move of the running loop.  Every game that does finish is still a complete, legal game from the empty board."""
    max_steps = 30
    eng = Engine('go', 9, num_games=10, max_simulations=32, max_parallel=4, net=(1, 16, 16), precision='fp32', max_steps=max_steps, seed=5,
                 sample_ring=900, binding=emu)
    eng.set_weights(_dummy_weights(1, 16, 16, 17, 82, 81))
    with pytest.raises(Exception):
        eng.selfplay_restart([0])  # loop not begun
    eng.selfplay_begin(8, 4, warm_up_steps=4, check_resign_after_steps=8, resign_threshold=-1.0, disable_resign_ratio=1.0)
    ticks_per_move = (8 + 4 + 3) // 4
    uids, n_games, total = set(), 0, 0
    restarted_at = {}
    for mv in range(24):  # stagger: slot g restarts after (g % 8) * 3 moves
        eng.selfplay_tick(ticks_per_move)
        if mv % 3 == 0:
            sl = [g for g in range(10) if (g % 8) * 3 == mv]
            eng.selfplay_restart(sl)
            for g in sl:
                restarted_at[g] = mv
    with pytest.raises(ValueError):
        eng.selfplay_restart([1, 1])
    c0 = eng.counters()
    assert c0['errors'] == 0
    # the slots now differ in age: the per-slot ply counters say so
    steps = [eng.env_scalars(g)['steps'] for g in range(10)]
    assert len(set(steps)) >= 5, steps
    # more simulations per move from here on
    eng.selfplay_update(4, 8, -1.0, 1.0, search=dict(num_simulations=24, num_parallel=4))
    for rnd in range(40):
        eng.selfplay_tick(7)
        games, states, pis, zs = eng.drain_games()
        for rec in games:
            total += _check_game('go', rec, states, pis, zs, eng.last_moves, max_steps)
            assert rec['reserved'] not in uids
            uids.add(rec['reserved'])
            n_games += 1
    c1 = eng.counters()
    assert c1['errors'] == 0 and c1['ring_dropped'] == 0
    assert c1['games'] == n_games and c1['samples'] == total and n_games >= 10
    sims_per_move = (c1['simulations'] - c0['simulations']) / max(1, c1['moves'] - c0['moves'])
    assert 20.0 < sims_per_move < 30.0, sims_per_move  # 24 + num_parallel bound, minus carried visits
    eng.close()


def test_bench_population_stagger(emu):
    """bench.py's prologue: after it the slots hold games of every age 0..L-1 (or younger, where a game ended on its own), nothing is
    left in the rings, and the loop runs on with the workload's search parameters."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    G, L = 24, 12
    eng = Engine('go', 9, num_games=G, max_simulations=32, max_parallel=4, net=(1, 16, 16), precision='fp32', max_steps=40, seed=9,
                 sample_ring=2000, binding=emu)
    eng.set_weights(_dummy_weights(1, 16, 16, 17, 82, 81))
    eng.selfplay_begin(4, 4, warm_up_steps=4, check_resign_after_steps=8, resign_threshold=-1.0, disable_resign_ratio=1.0)
    bench.stagger_population(eng, G, L, 4, 4, 8, 16)
    steps = np.array([eng.env_scalars(g)['steps'] for g in range(G)])
    want = L - 1 - (np.arange(G) % L)  # slot g restarted after (g mod L) + 1 of the L prologue moves
    assert np.all(steps <= want + 1) and np.all(steps[want > 2] > 0), (steps, want)
    assert len(set(steps.tolist())) >= L // 2
    assert eng.drain_games()[0] == []
    c0 = eng.counters()
    assert eng.tick_profile(True) == ({'select': 0.0, 'network': 0.0, 'expand_backup': 0.0, 'move_reroot': 0.0, 'tick': 0.0}, 0)
    eng.selfplay_tick(40)
    assert eng.tick_profile(False)[1] == 40  # the emulation build has no device clock: it only counts the profiled ticks
    c1 = eng.counters()
    spm = (c1['simulations'] - c0['simulations']) / max(1, c1['moves'] - c0['moves'])
    assert c1['errors'] == 0 and 12.0 < spm < 21.0, spm
    eng.close()


def test_drain_in_pieces_equals_drain_at_once(emu):
    """az_drain_games copies the samples of consecutive games as contiguous runs of the ring; draining game by game, with a tight
    sample budget, or all at once must hand out the same games with the same samples (small ring: the runs wrap)."""
    def play(drain):
        eng = Engine('go', 9, num_games=8, max_simulations=16, max_parallel=4, net=(1, 16, 16), precision='fp32', max_steps=24, seed=17,
                     sample_ring=600, binding=emu)
        eng.set_weights(_dummy_weights(1, 16, 16, 17, 82, 81))
        eng.selfplay_begin(12, 4, warm_up_steps=6, check_resign_after_steps=8, resign_threshold=-1.0, disable_resign_ratio=1.0)
        out = []
        for rnd in range(30):
            eng.selfplay_tick(12)
            while True:
                games, states, pis, zs = drain(eng)
                if not games:
                    break
                mv = eng.last_moves
                for g in games:
                    s0, ln = g['first_sample'], g['game_length']
                    out.append((g['reserved'], g['slot'], ln, g['winner'], states[s0:s0 + ln].tobytes(), pis[s0:s0 + ln].tobytes(), zs[s0:s0 + ln].tobytes(),
                                mv[s0:s0 + ln].tobytes()))
        c = eng.counters()
        eng.close()
        assert c['ring_dropped'] == 0 and c['games'] == len(out) and c['samples'] == sum(o[2] for o in out)
        return out

    whole = play(lambda e: e.drain_games())
    assert len(whole) > 16
    assert play(lambda e: e.drain_games(max_games=1)) == whole
    assert play(lambda e: e.drain_games(max_games=3, max_samples=60)) == whole
