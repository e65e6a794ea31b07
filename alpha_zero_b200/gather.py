"""All-gather of the (state, pi, z) samples produced by the game shards.

Games are independent, so the only exchange between the per-GPU processes is this one collective per collection round.
The samples never visit the host on the way: `az_gather_pack` packs the finished games of a rank device-to-device into
the send block, the ranks all-gather their COUNTS (one int64 each), and then exactly max(count) rows per rank travel —
`all_gather_into_tensor` straight on the device buffers, NCCL over NVLink on the GPU box, gloo in the CPU tests (where the
"device" of the host-emulation engine is host memory, so the same code runs).  The result stays on the device
(`DeviceReplay` / a learner on that GPU consumes it there); `to_host()` is for the rank where a host-side learner lives.

The reference's transport for the same records is `data_queue.put((game_seq, stats))` (core/pipeline.py:283,
training_go.py:279): pickled numpy arrays through an mp.Queue, one actor process per game.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_slots(total_games, rank, world):
    """Contiguous slice of game slots owned by `rank`: [rank*G/R, (rank+1)*G/R)."""
    lo = rank * total_games // world
    hi = (rank + 1) * total_games // world
    return lo, hi


class GatheredSamples:
    """Every rank's samples of one round, rank-major, on the gathering device."""

    def __init__(self, states, pis, zs, counts, shape, pinned=None):
        self.states, self.pis, self.zs, self.counts, self._shape, self._pinned = states, pis, zs, counts, shape, pinned

    def __len__(self):
        return int(self.zs.shape[0])

    def to_host(self):
        """numpy arrays (states int8 [n, planes, N, N], pis f32 [n, A], z f32 [n]).  On a CUDA device the three copies go
        asynchronously into the gatherer's page-locked buffers (one synchronisation; the arrays are views that the next
        to_host() overwrites — copy what must outlive the round); on the CPU (gloo tests) they are plain copies."""
        n = len(self)
        if self._pinned is None or not self.states.is_cuda:
            return (self.states.cpu().numpy().reshape((n,) + self._shape), self.pis.cpu().numpy(), self.zs.cpu().numpy())
        hs, hp, hz = self._pinned(n)
        hs[:n].copy_(self.states, non_blocking=True)
        hp[:n].copy_(self.pis, non_blocking=True)
        hz[:n].copy_(self.zs, non_blocking=True)
        torch.cuda.current_stream(self.states.device).synchronize()
        return (hs[:n].numpy().reshape((n,) + self._shape), hp[:n].numpy(), hz[:n].numpy())


class DeviceSampleGatherer:
    """One all-gather per collection round.  `capacity` bounds the samples a rank contributes per round; finished games beyond
    it stay in the engine's sample ring and travel with the next round (az_gather_pack stops at the last whole game that fits)."""

    def __init__(self, engine, capacity, device=None, group=None):
        self.engine, self.capacity, self.group = engine, int(capacity), group
        self.device = torch.device(device) if device is not None else torch.device('cpu')
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.sdim, self.adim = engine.obs_bytes, engine.A
        self.shape = (engine.planes, engine.N, engine.N)
        kw = dict(device=self.device)
        self.blk_s = torch.zeros((self.capacity, self.sdim), dtype=torch.int8, **kw)
        self.blk_p = torch.zeros((self.capacity, self.adim), dtype=torch.float32, **kw)
        self.blk_z = torch.zeros((self.capacity,), dtype=torch.float32, **kw)
        if self.world > 1:
            self.out_s = torch.empty((self.world * self.capacity, self.sdim), dtype=torch.int8, **kw)
            self.out_p = torch.empty((self.world * self.capacity, self.adim), dtype=torch.float32, **kw)
            self.out_z = torch.empty((self.world * self.capacity,), dtype=torch.float32, **kw)
            self.cnt = torch.zeros((1,), dtype=torch.int64, **kw)
            self.cnts = torch.zeros((self.world,), dtype=torch.int64, **kw)
        self.total = 0
        self.bytes_sent = 0
        self._host = None

    def _pinned(self, n):
        """Page-locked host staging for to_host(), grown on demand (only the rank that reads the samples ever allocates it)."""
        if self._host is None or self._host[0].shape[0] < n:
            rows = max(n, self.capacity)
            self._host = (torch.empty((rows, self.sdim), dtype=torch.int8, pin_memory=True),
                          torch.empty((rows, self.adim), dtype=torch.float32, pin_memory=True),
                          torch.empty((rows,), dtype=torch.float32, pin_memory=True))
        return self._host

    def push(self):
        """Pack this rank's finished games and exchange.  Returns (records of THIS rank's games, GatheredSamples of all ranks)."""
        games, n = self.engine.gather_pack(self.blk_s.data_ptr(), self.blk_p.data_ptr(), self.blk_z.data_ptr(), self.capacity)
        if self.world == 1:
            out = GatheredSamples(self.blk_s[:n], self.blk_p[:n], self.blk_z[:n], [n], self.shape, self._pinned)
            self.total += n
            return games, out
        self.cnt[0] = n
        dist.all_gather_into_tensor(self.cnts, self.cnt, group=self.group)
        counts = [int(c) for c in self.cnts.tolist()]  # world integers: the only host read of the exchange
        m = max(counts)
        if m == 0:
            return games, GatheredSamples(self.blk_s[:0], self.blk_p[:0], self.blk_z[:0], counts, self.shape, self._pinned)
        w = self.world
        dist.all_gather_into_tensor(self.out_s[: w * m], self.blk_s[:m], group=self.group)
        dist.all_gather_into_tensor(self.out_p[: w * m], self.blk_p[:m], group=self.group)
        dist.all_gather_into_tensor(self.out_z[: w * m], self.blk_z[:m], group=self.group)
        self.bytes_sent += m * (self.sdim + 4 * self.adim + 4)
        if all(c == m for c in counts):
            S, P, Z = self.out_s[: w * m], self.out_p[: w * m], self.out_z[: w * m]
        else:  # ragged: drop each rank's padding rows, on the device
            S = torch.cat([self.out_s[r * m: r * m + c] for r, c in enumerate(counts)])
            P = torch.cat([self.out_p[r * m: r * m + c] for r, c in enumerate(counts)])
            Z = torch.cat([self.out_z[r * m: r * m + c] for r, c in enumerate(counts)])
        self.total += sum(counts)
        return games, GatheredSamples(S, P, Z, counts, self.shape, self._pinned)
