"""TEST INFRASTRUCTURE ONLY — host emulation build of the engine's game / tree logic.

Compiles alpha_zero_b200/csrc/az_engine.cu with -DAZ_EMU (a "warp" of width 1, device memory = host
memory, kernels = loops; see csrc/az_warp.cuh) into tests/emu/libaz_emu.so so that the `-m "not gpu"`
tests can check the per-game device routines and the host-side C ABI logic against the oracle in the
build container, which has no GPU.  The package never loads this library: alpha_zero_b200._lib only
opens alpha_zero_b200/libaz_b200.so and fails loudly when CUDA is missing.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'alpha_zero_b200', 'csrc')
OUT = os.path.join(HERE, 'libaz_emu.so')
# 'deep': recorded path length 3, so that ordinary searches take the deeper-than-AZ_PATH fallbacks (serial parent walks for virtual
# loss / backup / leaf history), and every incremental group relabelling is cross-checked against a full one (aborts on a mismatch)
VARIANTS = {'': [], 'deep': ['-DAZ_PATH=3', '-DAZ_CHECK_LABELS']}


def build(force=False, variant=''):
    out = OUT if not variant else os.path.join(HERE, f'libaz_emu_{variant}.so')
    srcs = [os.path.join(CSRC, 'az_engine.cu'), os.path.join(HERE, 'az_emu_net_stub.cpp')]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))] + [os.path.abspath(__file__)]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(p) for p in deps):
        return out
    cmd = ['g++', '-std=c++17', '-O2', '-g', '-DAZ_EMU', '-ffp-contract=off', '-fPIC', '-shared', '-Wall', '-Wno-unused-function',
           '-Wno-unused-variable', '-Wno-unknown-pragmas'] + VARIANTS[variant] + ['-I', CSRC, '-I', os.path.join(ROOT, 'include'), '-x', 'c++', srcs[0], srcs[1], '-o', out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == '__main__':
    print(build(force=True))
