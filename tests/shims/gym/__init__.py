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
