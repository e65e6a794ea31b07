"""Reference-shaped Python API on the host-emulation binding (no GPU): env contract, seeded search traces, full games."""
import ctypes
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
import facadecheck  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402


@pytest.fixture(scope='module', autouse=True)
def emu_binding():
    facadecheck.use_binding(Binding(ctypes.CDLL(build_emu.build())))
    yield
    facadecheck.use_binding(None)


def test_env_contract():
    facadecheck.env_contract()


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_mcts_api_traces(game):
    assert facadecheck.mcts_api_traces(game) > 40


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_pipeline_traces(game):
    facadecheck.pipeline_traces(game)


def test_error_paths():
    facadecheck.error_paths()


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_batched_matches(game):
    facadecheck.batched_matches(game)


@pytest.mark.parametrize('game', ['go9', 'gomoku13'])
def test_device_matches(game):
    from alpha_zero_b200.envs import _pool

    facadecheck.device_matches(game, binding=_pool._TEST_BINDING)
