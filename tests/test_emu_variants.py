"""Equivalence of the engine's internal variants on the host-emulation build (no GPU needed):

* per-node position cache (AZ_NODE_CACHE, csrc/az_tree.cuh game_collect_nc) vs replaying every descent from the root:
  identical self-play trajectories (the stub evaluator hashes the leaf observation, so one differing observation byte,
  legal mask or visit count changes every later move), and the reference's MCTS traces / env corpus in both modes;
* the 'deep' build (recorded path length 3 + group-label cross-check): the deeper-than-AZ_PATH fallbacks give the same
  trajectories and pass the reference traces too.
"""
import ctypes
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
import enginecheck  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402
from alpha_zero_b200.engine import Engine  # noqa: E402
from test_emu_selfplay import _dummy_weights  # noqa: E402


@pytest.fixture(scope='module')
def libs():
    return {v: Binding(ctypes.CDLL(build_emu.build(variant=v))) for v in ('', 'deep')}


@pytest.fixture
def node_cache_env():
    old = {k: os.environ.get(k) for k in ('AZ_NODE_CACHE', 'AZ_EMU_SHARP', 'AZ_EMU_VALUE_SCALE')}
    os.environ['AZ_EMU_SHARP'] = '60'  # peaked stub policy + small values: trees several plies deep, well past the deep build's recorded path
    os.environ['AZ_EMU_VALUE_SCALE'] = '0.05'
    yield lambda v: os.environ.__setitem__('AZ_NODE_CACHE', str(v))
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _trajectory(binding, game, n, sims, par, ticks, num_stack=8):
    A = n * n + (1 if game == 'go' else 0)
    eng = Engine(game, n, num_games=6, max_simulations=sims, max_parallel=par, net=(1, 16, 16), precision='fp32', seed=21, max_steps=60 if game == 'go' else 0,
                 num_stack=num_stack, binding=binding)
    eng.set_weights(_dummy_weights(1, 16, 16, 2 * num_stack + 1, A, n * n if game == 'go' else (n + 4) ** 2))
    eng.selfplay_begin(sims, par, warm_up_steps=6, check_resign_after_steps=10, resign_threshold=-0.035, disable_resign_ratio=0.5)
    h = hashlib.sha1()
    for _ in range(ticks // 10):
        eng.selfplay_tick(10)
        games, states, pis, zs = eng.drain_games()
        for arr in (states, pis, zs, eng.last_moves):
            h.update(np.ascontiguousarray(arr).tobytes())
        h.update(repr([sorted(g.items()) for g in games]).encode())
    for g in range(6):
        h.update(eng.env_board(g).tobytes())
        h.update(eng.env_observation(g).tobytes())
    c = eng.counters()
    eng.close()
    assert c['errors'] == 0 and c['games'] > 0
    return h.hexdigest(), {k: c[k] for k in ('simulations', 'evaluations', 'moves', 'games', 'nodes', 'depth_sum', 'descents', 'samples')}


@pytest.mark.parametrize('case', [('go', 9, 96, 4, 8), ('go', 9, 40, 1, 8), ('gomoku', 9, 64, 4, 8), ('go', 5, 64, 8, 3), ('go', 19, 48, 4, 8), ('gomoku', 13, 48, 8, 8)])
def test_node_cache_and_deep_paths_give_identical_selfplay(libs, node_cache_env, case):
    game, n, sims, par, num_stack = case
    out = {}
    for variant in ('', 'deep'):
        for nc in (0, 1):
            node_cache_env(nc)
            out[(variant, nc)] = _trajectory(libs[variant], game, n, sims, par, 400, num_stack)
    base = out[('', 0)]
    for k, v in out.items():
        assert v == base, (k, v, base)
    if n <= 9:  # mean leaf depth: a good share of the descents passes the deep build's 3 recorded plies (wide boards search broadly)
        assert base[1]['depth_sum'] / base[1]['descents'] > 2.2


@pytest.mark.parametrize('variant,nc', [('', 1), ('deep', 0), ('deep', 1)])
@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_reference_traces_in_every_variant(libs, node_cache_env, variant, nc, game):
    node_cache_env(nc)
    assert enginecheck.mcts_traces(libs[variant], game) > 50
    enginecheck.concurrent_searches(libs[variant], game)


def _games_by_uid(binding, pipeline, game, ticks_per_call):
    """Self-play on 10 slots; every finished game keyed by its uid -> digest of (record, states, pi, z, moves)."""
    os.environ['AZ_PIPELINE'] = str(int(pipeline))
    try:
        n, A = 9, (82 if game == 'go' else 81)
        eng = Engine(game, n, num_games=10, max_simulations=32, max_parallel=4, net=(1, 16, 16), precision='fp32', seed=33, max_steps=50 if game == 'go' else 0,
                     binding=binding)
    finally:
        os.environ.pop('AZ_PIPELINE', None)
    eng.set_weights(_dummy_weights(1, 16, 16, 17, A, 81 if game == 'go' else 169))
    eng.selfplay_begin(24, 4, warm_up_steps=6, check_resign_after_steps=10, resign_threshold=-0.035, disable_resign_ratio=0.5)
    out = {}
    done = 0
    while done < 420:
        eng.selfplay_tick(ticks_per_call)
        done += ticks_per_call
        games, states, pis, zs = eng.drain_games()
        mv = eng.last_moves
        for g in games:
            s0, ln = g['first_sample'], g['game_length']
            h = hashlib.sha1()
            for arr in (states[s0:s0 + ln], pis[s0:s0 + ln], zs[s0:s0 + ln], mv[s0:s0 + ln]):
                h.update(np.ascontiguousarray(arr).tobytes())
            h.update(repr(sorted((k, v) for k, v in g.items() if k != 'first_sample')).encode())
            assert g['reserved'] not in out
            out[g['reserved']] = h.hexdigest()
    final = hashlib.sha1()
    for g in range(10):
        final.update(eng.env_board(g).tobytes())
        final.update(repr(sorted(eng.env_scalars(g).items())).encode())
    c = eng.counters()
    eng.close()
    assert c['errors'] == 0
    return out, final.hexdigest(), {k: c[k] for k in ('simulations', 'evaluations', 'moves', 'games', 'nodes', 'depth_sum', 'descents', 'samples')}


@pytest.mark.parametrize('game', ['go', 'gomoku'])
def test_two_half_pipeline_plays_the_same_games(libs, node_cache_env, game):
    """AZ_PIPELINE=1 (tree kernels of one half of the slots overlapped with the network of the other half) runs the same n leaf
    batches per call for every game: every finished game, the final positions and all counters equal the serial tick's, for
    different call granularities (the order in which finished games reach the ring may differ, hence the comparison by game id)."""
    node_cache_env(1)
    ref = _games_by_uid(libs[''], False, game, 7)
    assert len(ref[0]) >= 8
    for mode in (1, 2):  # 2 = persistent fused tree pass: expand/backup -> move -> collection of a game back to back in one warp
        for ticks_per_call in (7, 1, 20):
            got = _games_by_uid(libs[''], mode, game, ticks_per_call)
            assert got == ref, (mode, ticks_per_call, got[2], ref[2])
