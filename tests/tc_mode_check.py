"""Numerics of one tensor-core tower variant (AZ_TC_MODE) against the oracle's bf16-rounding torch restatement, on random positions.

    AZ_TC_MODE=5 python tests/tc_mode_check.py [go9_c2|gomoku13_c4|go9_small64] [n_leaves]

Test infrastructure (imports oracle/): used on the GPU box to validate an experimental kernel variant before it may become
the default; the default variant is covered by tests/test_gpu_engine.py.  Exit code 0 = within the bf16 tolerance (1e-2 on pi)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {'go9_c2': ('go', 9, 10, 128, 128), 'gomoku13_c4': ('gomoku', 13, 6, 64, 64), 'go9_small64': ('go', 9, 2, 64, 64)}


def main():
    from alpha_zero_b200.engine import Engine
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from oracle import net as onet

    tag = sys.argv[1] if len(sys.argv) > 1 else 'go9_c2'
    n_leaves = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    game, n, nb, nf, fc = CASES[tag]
    a = n * n + (1 if game == 'go' else 0)
    torch.manual_seed(5)
    net = randomize_batchnorm(AlphaZeroNet((17, n, n), a, nb, nf, fc, game == 'gomoku')).eval()
    rng = np.random.default_rng(3)
    x = (rng.random((n_leaves, 17, n, n)) < 0.3).astype(np.int8)
    x[:, 16] = (np.arange(n_leaves) % 2)[:, None, None]
    eng = Engine(game, n, num_games=37, max_simulations=8, max_parallel=4, net=(nb, nf, fc), precision='bf16')  # 148-leaf chunks: odd tile counts
    eng.set_weights(net.state_dict())
    pi, v = eng.net_forward(x)
    pi2, v2 = eng.net_forward(x[::-1].copy())  # second call on the same buffers: stale rows of the first call must not leak
    eng.close()
    lg, ve = onet.forward_bf16_emulated(net.state_dict(), torch.from_numpy(x).float(), game == 'gomoku')
    pe = torch.softmax(lg, dim=-1).numpy()
    e_pi, e_v = float(np.abs(pi - pe).max()), float(np.abs(v - ve.numpy()[:, 0]).max())
    e_pi2, e_v2 = float(np.abs(pi2[::-1] - pe).max()), float(np.abs(v2[::-1] - ve.numpy()[:, 0]).max())
    ok = e_pi < 1e-2 and e_v < 2e-2 and e_pi2 < 1e-2 and e_v2 < 2e-2 and bool(np.isfinite(pi).all())
    print(f'AZ_TC_MODE={os.environ.get("AZ_TC_MODE", "default")} {tag} leaves={n_leaves}: max|pi-emu|={e_pi:.3e} max|v-emu|={e_v:.3e} '
          f'(second pass {e_pi2:.3e} / {e_v2:.3e}) -> {"OK" if ok else "MISMATCH"}')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
