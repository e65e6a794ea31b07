"""Evaluation-dataset on-ramp on the CUDA library (B200): az_env_replay / eval_dataset.py against the reference's replay_sgf
goldens, and eval_on_pro_games (core/pipeline.py:868-941) with the forward pass on the engine's fp32 tower against the same
statistics computed from the oracle's torch forward."""
import numpy as np
import pytest
import torch

import datasetcheck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def cuda():
    from alpha_zero_b200 import _lib

    return _lib.load()


def test_replay_matches_reference(cuda):
    datasetcheck.replay_matches_reference(cuda)


def test_replay_abi_edges(cuda):
    datasetcheck.replay_abi_edges(cuda)


def test_eval_on_pro_games(cuda, tmp_path):
    from torch.utils.data import DataLoader

    from alpha_zero_b200 import eval_dataset as ed
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from oracle import net as onet

    ds = datasetcheck.dataset_and_metrics(cuda, tmp_path)
    torch.manual_seed(11)
    net = randomize_batchnorm(AlphaZeroNet((17, 9, 9), 82, 2, 32, 32, False)).eval()
    stats = ed.eval_on_pro_games(net, torch.device('cuda:0'), DataLoader(ds, batch_size=1024, shuffle=False), precision='fp32')
    states, target_pi, target_v = ds.tensors
    lg, v = onet.forward(net.state_dict(), states, False)
    p = torch.softmax(lg, dim=-1)
    n = len(states)
    want_entropy = float(-(p * torch.log(p)).sum(dim=1).mean())
    want_mse = float(((v[:, 0] - target_v) ** 2).mean())
    assert abs(stats['policy_entropy'] - want_entropy) < 1e-3 and abs(stats['value_mse_error'] - want_mse) < 1e-3
    _, pred = torch.topk(p, 5, dim=1)
    hit = pred.eq(torch.argmax(target_pi, dim=1).unsqueeze(1))
    for k in (1, 3, 5):
        want = float(hit[:, :k].any(dim=1).sum()) / n
        assert abs(stats[f'policy_top_{k}_accuracy'] - want) <= 3.0 / n, (k, stats, want)  # near-ties may flip under 1e-4 noise
