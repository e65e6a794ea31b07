"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/az_engine.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    from alpha_zero_b200 import build

    return build.build()


def test_header_symbols_exported(lib_path):
    from alpha_zero_b200 import _lib

    header = open(os.path.join(ROOT, 'include', 'az_engine.h')).read()
    declared = set(re.findall(r'\b(az_[a-z0-9_]+)\s*\(', header))
    dll = ctypes.CDLL(lib_path)
    missing = [s for s in declared if not hasattr(dll, s)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert dll.az_version() >= 100


def test_no_cpu_fallback(lib_path):
    """Without a CUDA device az_create must fail loudly (this container has no GPU)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from alpha_zero_b200.engine import Engine

    with pytest.raises(Exception) as ei:
        Engine('go', 9, num_games=1, max_simulations=8, max_parallel=1)
    assert 'CUDA' in str(ei.value) or 'cuda' in str(ei.value)


def test_sass_is_sm100a(lib_path):
    out = os.popen(f'cuobjdump -lelf {lib_path} 2>/dev/null').read()
    assert 'sm_100a' in out, out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'alpha_zero_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, re.M), f
                assert 'libaz_emu' not in text or f == '_lib.py', f
