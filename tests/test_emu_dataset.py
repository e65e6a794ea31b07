"""Evaluation-dataset on-ramp on the host emulation build of the device code (no GPU): az_env_replay + eval_dataset.py against
the reference's replay_sgf goldens."""
import ctypes
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
import datasetcheck  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402


@pytest.fixture(scope='module')
def binding():
    return Binding(ctypes.CDLL(build_emu.build()))


def test_replay_matches_reference(binding):
    datasetcheck.replay_matches_reference(binding)


def test_replay_abi_edges(binding):
    datasetcheck.replay_abi_edges(binding)


def test_build_eval_dataset(binding, tmp_path):
    datasetcheck.dataset_and_metrics(binding, tmp_path)
