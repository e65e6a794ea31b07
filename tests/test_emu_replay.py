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
    def make_out(batch, obs_bytes, A):  # the emulation's "device" memory is host memory
        import numpy as np

        st, pi, z = np.zeros((batch, obs_bytes), np.int8), np.zeros((batch, A), np.float32), np.zeros(batch, np.float32)
        return (st.ctypes.data, pi.ctypes.data, z.ctypes.data), lambda: (st, pi, z)

    replaycheck.uniform_replay_semantics(emu, make_out)


def test_ingest_equals_drain(emu):
    replaycheck.ingest_equals_drain(emu)
