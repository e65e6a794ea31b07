"""World-size-2 gloo run of the only multi-process step of the path: the sample all-gather + slot sharding (CPU)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    import torch.distributed as dist

    from alpha_zero_b200.gather import all_gather_samples, shard_slots

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_slots(4096, rank, world)
    rng = np.random.RandomState(100 + rank)
    n = 5 + 7 * rank  # ragged: ranks hold different numbers of samples, rank 0 more than zero, one rank could be empty
    st = rng.randint(0, 2, size=(n, 17, 9, 9)).astype(np.int8)
    pi = rng.rand(n, 82).astype(np.float32)
    z = rng.choice([-1.0, 0.0, 1.0], size=n).astype(np.float32)
    S, P, Z, kept = all_gather_samples(st, pi, z, capacity=16)
    S2, P2, Z2, kept2 = all_gather_samples(st[:0], pi[:0], z[:0], capacity=4)  # empty contribution from every rank
    # bench.py's exchange: a block of 6 samples per rank and round, the rest waits; three rounds, the last two with nothing new
    from alpha_zero_b200.gather import SampleGatherer

    gt = SampleGatherer(capacity=6)
    rounds = [gt.push(st, pi, z)] + [gt.push(st[:0], pi[:0], z[:0]) for _ in range(2)]
    carried = ([r[2].tolist() for r in rounds], [r[0].shape for r in rounds], gt.pending(), gt.total)
    out.put((rank, lo, hi, S.shape, float(S.sum()), float(P.sum()), Z.tolist(), kept, S2.shape[0], float(st.sum()), float(pi.sum()), z.tolist(), carried))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_samples_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, r1) = res
    assert (r0[1], r0[2], r1[1], r1[2]) == (0, 2048, 2048, 4096)
    for r in res:
        assert r[3] == (5 + 12, 17, 9, 9) and r[7] == 0 and r[8] == 0
        assert abs(r[4] - (r0[9] + r1[9])) < 1e-3 and abs(r[5] - (r0[10] + r1[10])) < 1e-2
        assert r[6] == r0[11] + r1[11]  # rank order preserved
        # rank 0 holds 5 samples, rank 1 holds 12: round 1 moves 5 + 6, round 2 the next 6 of rank 1, round 3 nothing; order kept
        zs_rounds, shapes, pending, total = r[12]
        assert zs_rounds[0] == r0[11] + r1[11][:6] and zs_rounds[1] == r1[11][6:] and zs_rounds[2] == []
        assert shapes[0] == (11, 17, 9, 9) and shapes[1] == (6, 17, 9, 9) and pending == 0 and total == 17
