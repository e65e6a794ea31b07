"""Host-emulation build of the device code (tests/emu) vs goldens — runs without a GPU.

This does NOT exercise the product library; it checks the per-game routines of csrc/az_board.cuh and
csrc/az_tree.cuh plus the host side of the C ABI before any GPU time is spent.  The GPU tests
(tests/test_gpu_*.py) run the same checks on libaz_b200.so.
"""
import ctypes
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
import enginecheck  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402


@pytest.fixture(scope='module')
def emu():
    return Binding(ctypes.CDLL(build_emu.build()))


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_env_corpus(emu, game):
    bad, n = enginecheck.replay_corpus(emu, game, stride=int(os.environ.get('AZ_CORPUS_STRIDE', '12')))
    assert not bad, f'{len(bad)}/{n} games differ, first {bad[:5]}'


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_env_observation_copy_export(emu, game):
    enginecheck.final_observations(emu, game, count=8)


def test_go19_unit(emu):
    enginecheck.go19_unit(emu)


def test_gomoku_unit(emu):
    enginecheck.gomoku_unit(emu)


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_mcts_traces(emu, game):
    assert enginecheck.mcts_traces(emu, game) > 50


@pytest.mark.parametrize('name', sorted(enginecheck.EXTRA))
def test_extra_corpora(emu, name):
    bad, n = enginecheck.replay_extra(emu, name, stride=1 if name != 'pro_go9' else int(os.environ.get('AZ_CORPUS_STRIDE', '10')))
    assert not bad, f'{name}: {len(bad)}/{n} games differ, first {bad[:5]}'


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_concurrent_searches(emu, game):
    enginecheck.concurrent_searches(emu, game)
