"""World-size-2 gloo run of the only multi-process step of the path: slot sharding + the device-resident sample all-gather
(alpha_zero_b200/gather.py) over two self-play engines.  CPU only: the engines are the host-emulation build (whose "device"
memory is host memory, so az_gather_pack fills CPU torch tensors) and the collective is gloo; on the GPU box the same code
runs on libaz_b200.so + NCCL (bench.py --gpus N)."""
import ctypes
import hashlib
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _worker(rank, world, port, out):
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(HERE, 'emu'))
    import build_emu
    from alpha_zero_b200._lib import Binding
    from alpha_zero_b200.engine import Engine
    from alpha_zero_b200.gather import DeviceSampleGatherer, shard_slots
    from test_emu_selfplay import _dummy_weights

    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_slots(4096, rank, world)
    emu = Binding(ctypes.CDLL(build_emu.build()))

    def make():
        e = Engine('go', 9, num_games=6 + 4 * rank, max_simulations=16, max_parallel=4, net=(1, 16, 16), precision='fp32', max_steps=20 + 6 * rank,
                   seed=11 + rank, sample_ring=1200, binding=emu)
        e.set_weights(_dummy_weights(1, 16, 16, 17, 82, 81))
        e.selfplay_begin(8, 4, warm_up_steps=4, check_resign_after_steps=8, resign_threshold=-1.0, disable_resign_ratio=1.0)
        return e

    eng, twin = make(), make()  # equal seeds: the twin's host drain is what this rank's contribution must be
    cap = 100                    # samples per rank and round: less than what finishes, so games are carried over
    gt = DeviceSampleGatherer(eng, capacity=cap)
    mine, rounds = [], []
    for rnd in range(8):
        ticks = 30 if rnd < 5 else 0  # the last rounds only flush what is still queued
        if ticks:
            eng.selfplay_tick(ticks)
            twin.selfplay_tick(ticks)
            g2, st, pi, z = twin.drain_games(copy=True)
            mine.append((st, pi, z))
        games, got = gt.push()
        S, P, Z = got.to_host()
        assert len(got) == sum(got.counts) and got.counts[rank] == sum(g['game_length'] for g in games) <= cap
        off = sum(got.counts[:rank])
        rounds.append((got.counts, _digest(S, P, Z), [_digest(S[sum(got.counts[:r]):sum(got.counts[:r + 1])], P[sum(got.counts[:r]):sum(got.counts[:r + 1])],
                                                              Z[sum(got.counts[:r]):sum(got.counts[:r + 1])]) for r in range(world)],
                       (S[off:off + got.counts[rank]], P[off:off + got.counts[rank]], Z[off:off + got.counts[rank]])))
    # this rank's slices over all rounds == everything its twin drained on the host, in order
    own = [np.concatenate([r[3][k] for r in rounds]) for k in range(3)]
    exp = [np.concatenate([m[k] for m in mine]) for k in range(3)]
    same = all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(own, exp))
    c = eng.counters()
    out.put((rank, lo, hi, same, int(exp[2].shape[0]), [r[0] for r in rounds], [r[1] for r in rounds], [r[2] for r in rounds], c['errors'], c['ring_dropped'],
             gt.total))
    eng.close()
    twin.close()
    dist.barrier()
    dist.destroy_process_group()


def test_device_gather_world2():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, r1) = res
    assert (r0[1], r0[2], r1[1], r1[2]) == (0, 2048, 2048, 4096)
    assert r0[3] and r1[3]                      # every rank's own slice == its host drain, bit for bit, in order
    assert r0[4] > 100 and r1[4] > 100          # more than one block's worth finished: the carry-over path ran
    assert r0[5] == r1[5] and r0[6] == r1[6] and r0[7] == r1[7]   # both ranks saw the same counts and the same bytes, per rank segment
    assert any(c[0] != c[1] for c in r0[5])     # ragged rounds happened
    assert r0[5][-1] == [0, 0]                  # everything was flushed
    assert r0[8] == r1[8] == 0 and r0[9] == r1[9] == 0
    assert r0[10] == r1[10] == r0[4] + r1[4]    # nothing lost, nothing duplicated
