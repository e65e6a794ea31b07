"""Build alpha_zero_b200/libaz_b200.so in-tree with nvcc for sm_100a (no torch extension machinery).

    python -m alpha_zero_b200.build        # or __graft_entry__.build()
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, 'libaz_b200.so')
SOURCES = ['az_engine.cu', 'az_net.cu', 'az_net_tc.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--expt-relaxed-constexpr',
    '-Xcompiler', '-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function,-Wno-unused-variable', '-Xptxas', '-v',
    '-I', CSRC, '-I', os.path.join(ROOT, 'include'),
]


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))] + [os.path.join(ROOT, 'include', 'az_engine.h')]
    newest = max(os.path.getmtime(p) for p in deps)
    objs = []
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    for s in srcs:
        o = os.path.join(HERE, 'build', os.path.basename(s) + '.o')
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) >= newest:
            continue
        log = open(o + '.log', 'w')
        procs.append((s, subprocess.Popen([nvcc] + NVCC_FLAGS + ['-c', s, '-o', o], stdout=log, stderr=subprocess.STDOUT), o + '.log'))
    failed = False
    for s, p, logf in procs:
        rc = p.wait()
        if rc != 0 or verbose:
            sys.stderr.write(open(logf).read())
        failed = failed or rc != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if procs or force or not os.path.exists(OUT):
        subprocess.run([nvcc, '-shared', '-o', OUT] + objs + ['-lcudart'], check=True)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
