"""Device replay + fused augmentation on the host-emulation build: transformations bit-exact against the
reference's own outputs (tests/golden/transform.npz), circular storage and sampling like UniformReplay, ring -> replay ingestion."""
import ctypes
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
import build_emu  # noqa: E402
import replaycheck  # noqa: E402
from alpha_zero_b200._lib import Binding  # noqa: E402


@pytest.fixture(scope='module')
def emu():
    return Binding(ctypes.CDLL(build_emu.build()))


@pytest.mark.parametrize('tag', ['go9', 'gomoku13', 'go19'])
def test_transformations_match_reference(emu, tag):
    replaycheck.transformations(emu, tag)


def test_uniform_replay_semantics(emu):
    replaycheck.uniform_replay_semantics(emu)


def test_ingest_equals_drain(emu):
    replaycheck.ingest_equals_drain(emu)
