"""Checks of the device replay, shared by the emulation (CPU) and CUDA (GPU) test files."""
import os
import random

import numpy as np

from alpha_zero_b200.engine import Engine
from alpha_zero_b200.replay import TRANSFORMATIONS, DeviceReplay, Transition

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def transformations(binding, tag):
    z = np.load(os.path.join(GOLDEN, 'transform.npz'))
    kind, n = {'go9': ('go', 9), 'gomoku13': ('gomoku', 13), 'go19': ('go', 19)}[tag]
    eng = Engine(kind, n, num_games=1, max_simulations=8, max_parallel=1, binding=binding)
    eng.replay_create(16)
    st, pi, v = z[f'{tag}/state'], z[f'{tag}/pi'], z[f'{tag}/value']
    eng.replay_add(st, pi, v)
    idx = np.arange(len(v), dtype=np.int32)
    s0, p0, v0 = eng.replay_sample(idx, 0)
    np.testing.assert_array_equal(s0, st)
    np.testing.assert_array_equal(p0, pi)
    np.testing.assert_array_equal(v0, v)
    assert list(z['names']) == TRANSFORMATIONS
    for t, name in enumerate(TRANSFORMATIONS, start=1):
        s1, p1, v1 = eng.replay_sample(idx[::-1].copy(), t)
        np.testing.assert_array_equal(s1, z[f'{tag}/{name}/state'][::-1], err_msg=name)
        np.testing.assert_array_equal(p1, z[f'{tag}/{name}/pi'][::-1], err_msg=name)  # pure permutation: bit-exact
        np.testing.assert_array_equal(v1, v[::-1])
    eng.close()


def uniform_replay_semantics(binding, make_out=None):
    """Circular overwrite, counters and seeded sampling behave like core/replay.py:35-116 (checked against a plain-Python model).
    `make_out(batch, obs_bytes, A)` -> ((states_ptr, pis_ptr, values_ptr), read) supplies caller-owned "device" buffers for the
    zero-copy path: DeviceReplay.sample(out=...) must say True (None stays reserved for "not enough samples yet")."""
    eng = Engine('go', 9, num_games=1, max_simulations=8, max_parallel=1, binding=binding)
    rp = DeviceReplay(eng, capacity=10, random_state=np.random.RandomState(4))
    model = [None] * 10
    added = 0
    rng = np.random.RandomState(0)
    assert rp.sample(4) is None and rp.size == 0
    for g in range(5):
        ln = int(rng.randint(2, 6))
        seq = [Transition(state=(rng.rand(17, 9, 9) < 0.3).astype(np.int8), pi_prob=rng.rand(82).astype(np.float32), value=float(rng.choice([-1.0, 1.0])))
               for _ in range(ln)]
        rp.add_game(seq)
        for t in seq:
            model[added % 10] = t
            added += 1
    assert rp.num_games_added == 5 and rp.num_samples_added == added and rp.size == min(added, 10)
    ref_rng = np.random.RandomState(4)
    for _ in range(3):
        batch = rp.sample(6)
        idx = ref_rng.randint(low=0, high=10, size=6)
        np.testing.assert_array_equal(batch.state, np.stack([model[i].state for i in idx]))
        np.testing.assert_array_equal(batch.pi_prob, np.stack([model[i].pi_prob for i in idx]))
        np.testing.assert_array_equal(batch.value, np.array([model[i].value for i in idx], dtype=np.float32))
    if make_out is not None:
        ptrs, read = make_out(6, 17 * 81, 82)
        assert rp.sample(6, out=ptrs) is True
        idx = ref_rng.randint(low=0, high=10, size=6)
        st, pi, z = read()
        np.testing.assert_array_equal(st.reshape(6, 17, 9, 9), np.stack([model[i].state for i in idx]))
        np.testing.assert_array_equal(pi, np.stack([model[i].pi_prob for i in idx]))
        np.testing.assert_array_equal(z, np.array([model[i].value for i in idx], dtype=np.float32))
        assert DeviceReplay(Engine('go', 9, num_games=1, max_simulations=8, max_parallel=1, binding=binding), 4, np.random.RandomState(1)).sample(2, out=ptrs) is None
    # augmentation consumes python's `random` exactly like apply_random_transformation
    random.seed(9)
    expect = []
    for _ in range(8):
        expect.append(1 + TRANSFORMATIONS.index(random.choice(TRANSFORMATIONS)) if random.random() > 0.5 else 0)
    random.seed(9)
    seen = []
    real = eng.replay_sample

    def spy(indices, transform=0, out=None):
        seen.append(transform)
        return real(indices, transform, out=out)

    eng.replay_sample = spy
    for _ in range(8):
        rp.sample(4, augment=True)
    assert seen == expect and any(seen)
    try:
        eng.replay_sample = real
        real(np.array([11], dtype=np.int32), 0)
        raise AssertionError('out-of-range index accepted')
    except ValueError:
        pass
    eng.close()


def _dummy_weights(blocks, filters, fc, planes, A, hw):
    shapes = [(filters * planes * 9,)] + [(filters,)] * 4
    for _ in range(blocks * 2):
        shapes += [(filters * filters * 9,)] + [(filters,)] * 4
    shapes += [(2 * filters,)] + [(2,)] * 4 + [(A * 2 * hw,), (A,)] + [(filters,)] + [(1,)] * 4 + [(fc * hw,), (fc,), (fc,), (1,)]
    return {f't{i}': np.zeros(s, dtype=np.float32) for i, s in enumerate(shapes)}


def ingest_equals_drain(binding, weights=None, precision='fp32'):
    """Two engines with the same seed play the same games; one drains to the host, the other ingests ring -> replay on the
    device: the replay must then hold exactly the drained samples, in order (add_game per finished game)."""
    outs = []
    for mode in ('drain', 'ingest'):
        eng = Engine('go', 9, num_games=8, max_simulations=16, max_parallel=4, net=(1, 16, 16) if weights is None else weights[0], precision=precision,
                     max_steps=20, seed=7, sample_ring=400, binding=binding)
        eng.set_weights(_dummy_weights(1, 16, 16, 17, 82, 81) if weights is None else weights[1])
        eng.selfplay_begin(12, 4, warm_up_steps=4, check_resign_after_steps=50, resign_threshold=-1.0, disable_resign_ratio=1.0)
        S, P, Z = [], [], []
        if mode == 'ingest':
            eng.replay_create(4000)
        total = 0
        for _ in range(30):
            eng.selfplay_tick(6)
            if mode == 'drain':
                games, st, pi, z = eng.drain_games()
                S.append(st.copy()); P.append(pi.copy()); Z.append(z.copy())
            else:
                ng, ns = eng.replay_ingest()
                total += ns
        if mode == 'ingest':
            info = eng.replay_info()
            assert info['num_samples_added'] == total and info['size'] == total and info['num_games_added'] > 0
            s, p, z = eng.replay_sample(np.arange(total, dtype=np.int32), 0)
            outs.append((s, p, z))
        else:
            outs.append((np.concatenate(S), np.concatenate(P), np.concatenate(Z)))
        eng.close()
    # games that finish in the same tick reach the ring in the order their warps win the head atomic, which is not
    # deterministic on the GPU: compare the two sample sets independently of order
    def keyed(t):
        s, p, z = t
        keys = [s[i].tobytes() + p[i].tobytes() + z[i:i + 1].tobytes() for i in range(len(z))]
        return sorted(keys)

    assert len(outs[0][2]) == len(outs[1][2]) and len(outs[0][2]) > 50
    assert keyed(outs[0]) == keyed(outs[1])
