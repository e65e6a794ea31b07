"""Device replay + fused augmentation on the CUDA library (B200): same checks as tests/test_emu_replay.py, plus sampling
straight into the learner's CUDA tensors."""
import numpy as np
import pytest
import torch

import replaycheck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cuda():
    from alpha_zero_b200 import _lib

    return _lib.load()


@pytest.mark.parametrize('tag', ['go9', 'gomoku13', 'go19'])
def test_transformations_match_reference(cuda, tag):
    replaycheck.transformations(cuda, tag)


def test_uniform_replay_semantics(cuda):
    def make_out(batch, obs_bytes, A):
        st = torch.zeros((batch, obs_bytes), dtype=torch.int8, device='cuda')
        pi = torch.zeros((batch, A), dtype=torch.float32, device='cuda')
        z = torch.zeros((batch,), dtype=torch.float32, device='cuda')
        return (st.data_ptr(), pi.data_ptr(), z.data_ptr()), lambda: (st.cpu().numpy(), pi.cpu().numpy(), z.cpu().numpy())

    replaycheck.uniform_replay_semantics(cuda, make_out)


def test_ingest_equals_drain(cuda):
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm

    torch.manual_seed(3)
    net = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 82, 1, 64, 32, False)).eval()
    replaycheck.ingest_equals_drain(cuda, weights=((1, 64, 32), net.state_dict()), precision='bf16')


def test_sample_into_cuda_tensors(cuda):
    """outputs_on_device: the minibatch lands in torch CUDA tensors (what compute_losses consumes, pipeline.py:634-640)."""
    from alpha_zero_b200.engine import Engine

    eng = Engine('go', 9, num_games=1, max_simulations=8, max_parallel=1)
    eng.replay_create(64)
    rng = np.random.RandomState(1)
    st = (rng.rand(40, 17, 9, 9) < 0.3).astype(np.int8)
    pi = rng.rand(40, 82).astype(np.float32)
    z = rng.choice([-1.0, 1.0], size=40).astype(np.float32)
    eng.replay_add(st, pi, z)
    idx = rng.randint(0, 40, size=32).astype(np.int32)
    ts = torch.empty((32, 17, 9, 9), dtype=torch.int8, device='cuda')
    tp = torch.empty((32, 82), dtype=torch.float32, device='cuda')
    tz = torch.empty((32,), dtype=torch.float32, device='cuda')
    eng.replay_sample(idx, 4, out=(ts.data_ptr(), tp.data_ptr(), tz.data_ptr()))
    hs, hp, hz = eng.replay_sample(idx, 4)
    np.testing.assert_array_equal(ts.cpu().numpy(), hs)
    np.testing.assert_array_equal(tp.cpu().numpy(), hp)
    np.testing.assert_array_equal(tz.cpu().numpy(), hz)
    np.testing.assert_array_equal(hs, st[idx][:, :, ::-1, ::-1])  # rotate180
    eng.close()
