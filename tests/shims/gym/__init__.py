"""Minimal stand-in for the `gym` package (absent from this image).

Test infrastructure only: lets /root/reference import unmodified so it can act as
the oracle-of-oracles when golden vectors are generated (tests/golden/make_golden.py).
The reference reads nothing but `.shape` / `.n` from the spaces (envs/base.py:58-69).
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self, **kwargs):
        return None

    def close(self):
        return None


def _install_emulation_binding():
    """tools/run_reference_training_go.py --emu: every process of the reference's training driver imports `gym` (envs/base.py),
    so this is where the spawned actor process learns that its engine is the host-emulation build (tests/emu).  The product never
    does this: alpha_zero_b200._lib.load opens libaz_b200.so only."""
    import ctypes
    import os
    import sys

    emu_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'emu')
    if emu_dir not in sys.path:
        sys.path.insert(0, emu_dir)
    import build_emu
    from alpha_zero_b200 import _lib

    binding = _lib.Binding(ctypes.CDLL(build_emu.build()))
    _lib.load = lambda: binding


import os as _os

if _os.environ.get('AZ_TEST_EMU_BINDING') == '1':
    _install_emulation_binding()
