import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from alpha_zero_b200.engine import Engine
from alpha_zero_b200.network import AlphaZeroNet, randomize_batchnorm
z = np.load('tests/golden/net.npz')
def case(tag, prec):
    n, a, nb, nf, fc, gomoku = (int(v) for v in z[tag + '/cfg'])
    torch.manual_seed(123)
    net = randomize_batchnorm(AlphaZeroNet((17, n, n), a, nb, nf, fc, bool(gomoku))).eval()
    try:
        eng = Engine('gomoku' if gomoku else 'go', n, num_games=4, max_simulations=8, max_parallel=2, net=(nb, nf, fc), precision=prec)
        eng.set_weights(net.state_dict())
        pi, v = eng.net_forward(z[tag + '/x'])
        msg = ''
        if prec == 'bf16':
            from oracle import net as onet
            lg, ve = onet.forward_bf16_emulated(net.state_dict(), torch.from_numpy(z[tag + '/x']).float(), bool(gomoku))
            pe = torch.softmax(lg, -1).numpy()
            msg = f' | vs bf16 emulation: max|dpi| {float(np.abs(pi - pe).max()):.2e} max|dv| {float(np.abs(v - ve.numpy()[:, 0]).max()):.2e}'
        print(tag, prec, 'max|dpi|', float(np.abs(pi - z[tag + '/pi']).max()), 'max|dv|', float(np.abs(v - z[tag + '/v'][:, 0]).max()), msg, flush=True)
        eng.close()
    except Exception as ex:
        print(tag, prec, 'ERROR', repr(ex), flush=True)
def case19(prec):
    # 19x19, 256 filters (the C5 geometry, 2 blocks): no golden, compare against the oracle's torch forward / bf16 emulation
    from oracle import net as onet
    torch.manual_seed(7)
    net = randomize_batchnorm(AlphaZeroNet((17, 19, 19), 362, 2, 256, 256, False)).eval()
    x = (torch.rand((5, 17, 19, 19)) < 0.3).to(torch.int8).numpy()
    eng = Engine('go', 19, num_games=4, max_simulations=8, max_parallel=2, net=(2, 256, 256), precision=prec)
    eng.set_weights(net.state_dict())
    pi, v = eng.net_forward(x)
    lg, vr = onet.forward(net.state_dict(), torch.from_numpy(x).float(), False)
    pr = torch.softmax(lg, -1).numpy()
    msg = ''
    if prec == 'bf16':
        lg2, ve = onet.forward_bf16_emulated(net.state_dict(), torch.from_numpy(x).float(), False)
        pe = torch.softmax(lg2, -1).numpy()
        msg = f' | vs bf16 emulation: max|dpi| {float(np.abs(pi - pe).max()):.2e} max|dv| {float(np.abs(v - ve.numpy()[:, 0]).max()):.2e}'
    print('go19_256', prec, 'max|dpi|', float(np.abs(pi - pr).max()), 'max|dv|', float(np.abs(v - vr.numpy()[:, 0]).max()), msg, flush=True)
    eng.close()


which = sys.argv[1:] or ['fp32', 'bf16']
for prec in which:
    for tag in (['go9_small', 'gomoku13_small', 'go9_c2', 'gomoku13_c4'] if prec == 'fp32' else ['gomoku13_c4', 'go9_c2']):
        case(tag, prec)
    case19(prec)
