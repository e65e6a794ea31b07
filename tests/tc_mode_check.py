"""Numerics of the tensor-core tower variants (AZ_TC_MODE) against each other and against the oracle's bf16-rounding torch
restatement, on positions from random play (realistic, sparse planes; random dense planes saturate these random-init nets).

    python tests/tc_mode_check.py [go9_c2|gomoku13_c4|go9_small64] [n_positions] [modes, e.g. 4,5]

Test infrastructure (imports oracle/): used on the GPU box to validate an experimental kernel variant before it may become
the default; the default variant is covered by tests/test_gpu_engine.py.  1024 leaves per network call, so the persistent
kernels run several work units per CTA.  Exit code 0 = every listed mode is within the bf16 tolerance of the emulation (1e-2
on pi, 2e-2 on v) or within 1.5x + 2e-3 of the first mode's own distance to it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {'go9_c2': ('go', 9, 10, 128, 128), 'gomoku13_c4': ('gomoku', 13, 6, 64, 64), 'go9_small64': ('go', 9, 2, 64, 64),
         'go19_128': ('go', 19, 2, 128, 128), 'go13_128': ('go', 13, 3, 128, 64)}


def random_play_positions(game, n, count, seed):
    from oracle.boards import GoBoard, GomokuBoard

    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        env = GoBoard(n) if game == 'go' else GomokuBoard(n)
        obs = env.reset()
        while not env.is_game_over() and len(out) < count:
            out.append(np.asarray(obs, dtype=np.int8))
            legal = np.flatnonzero(np.asarray(env.legal_actions)[: n * n])  # no passes: keeps the boards filling up
            if legal.size == 0:
                break
            obs, _, _, _ = env.step(int(rng.choice(legal)))
    return np.stack(out)


def compare_modes(tag, count, modes):
    """{mode: dict(pi, v, max_dpi, max_dv, mean_dpi, repeat)} against the emulation, plus the emulation's mean top probability."""
    from alpha_zero_b200.engine import Engine
    from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
    from oracle import net as onet

    game, n, nb, nf, fc = CASES[tag]
    a = n * n + (1 if game == 'go' else 0)
    torch.manual_seed(123)
    net = randomize_batchnorm(AlphaZeroNet((17, n, n), a, nb, nf, fc, game == 'gomoku')).eval()
    x = random_play_positions(game, n, count, 3)
    lg, ve = onet.forward_bf16_emulated(net.state_dict(), torch.from_numpy(x).float(), game == 'gomoku')
    pe, ve = torch.softmax(lg, dim=-1).numpy(), ve.numpy()[:, 0]
    res = {}
    old = os.environ.get('AZ_TC_MODE')
    try:
        for m in modes:
            os.environ['AZ_TC_MODE'] = m
            eng = Engine(game, n, num_games=256, max_simulations=8, max_parallel=4, net=(nb, nf, fc), precision='bf16')
            eng.set_weights(net.state_dict())
            pi, v = eng.net_forward(x)
            pi2, v2 = eng.net_forward(x[::-1].copy())  # same buffers again: stale rows of the first call must not leak
            eng.close()
            res[m] = dict(pi=pi, v=v, max_dpi=float(np.abs(pi - pe).max()), max_dv=float(np.abs(v - ve).max()), mean_dpi=float(np.abs(pi - pe).mean()),
                          repeat=max(float(np.abs(pi2[::-1] - pi).max()), float(np.abs(v2[::-1] - v).max())), finite=bool(np.isfinite(pi).all()))
    finally:
        if old is None:
            os.environ.pop('AZ_TC_MODE', None)
        else:
            os.environ['AZ_TC_MODE'] = old
    return res, float(pe.max(axis=1).mean()), len(x)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'go9_c2'
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    modes = (sys.argv[3] if len(sys.argv) > 3 else os.environ.get('AZ_TC_MODE', '4')).split(',')
    res, top, npos = compare_modes(tag, count, modes)
    ok = True
    base = res[modes[0]]
    for m in modes:
        r = res[m]
        good = (r['max_dpi'] < 1e-2 and r['max_dv'] < 2e-2) or (r['max_dpi'] < 1.5 * base['max_dpi'] + 2e-3 and r['max_dv'] < 1.5 * base['max_dv'] + 2e-3)
        good = good and r['repeat'] < 1e-6 and r['finite'] and top < 0.999
        d_pi, d_v = float(np.abs(r['pi'] - base['pi']).max()), float(np.abs(r['v'] - base['v']).max())
        print(f'{tag} mode {m}: vs emulation max|dpi|={r["max_dpi"]:.3e} mean|dpi|={r["mean_dpi"]:.2e} max|dv|={r["max_dv"]:.3e}; vs mode {modes[0]} '
              f'max|dpi|={d_pi:.3e} max|dv|={d_v:.3e}; repeat call diff {r["repeat"]:.1e} -> {"OK" if good else "MISMATCH"}', flush=True)
        ok = ok and good
    print(f'{tag}: positions={npos} mean max-prob of the emulation {top:.3f} -> {"ALL OK" if ok else "FAILED"}')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
