"""All-gather of the (state, pi, z) samples produced by the game shards.

Games are independent, so the only exchange between the per-GPU processes is this one collective per
collection round: every rank contributes a fixed-capacity block [count | samples...] (padding keeps the
collective shape static) and receives everybody's.  Backend NCCL over NVLink on the GPU box (tensors on
the rank's device), gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_slots(total_games, rank, world):
    """Contiguous slice of game slots owned by `rank`: [rank*G/R, (rank+1)*G/R)."""
    lo = rank * total_games // world
    hi = (rank + 1) * total_games // world
    return lo, hi


def all_gather_samples(states, pis, zs, capacity, device=None, group=None):
    """states int8 [n, ...], pis float32 [n, A], zs float32 [n] (numpy) -> concatenated arrays of every rank, in rank order.
    Samples beyond `capacity` stay with the caller for the next round (returned as `kept`)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = min(len(zs), capacity)
    kept = len(zs) - n
    if world == 1:
        return states[:n], pis[:n], zs[:n], kept
    dev = torch.device(device) if device is not None else torch.device('cpu')
    sdim = int(np.prod(states.shape[1:])) if states.ndim > 1 else 1
    adim = pis.shape[1]
    blk_s = torch.zeros((capacity, sdim), dtype=torch.int8, device=dev)
    blk_p = torch.zeros((capacity, adim), dtype=torch.float32, device=dev)
    blk_z = torch.zeros((capacity + 1,), dtype=torch.float32, device=dev)
    if n:
        blk_s[:n].copy_(torch.from_numpy(np.ascontiguousarray(states[:n]).reshape(n, sdim)))
        blk_p[:n].copy_(torch.from_numpy(np.ascontiguousarray(pis[:n])))
        blk_z[:n].copy_(torch.from_numpy(np.ascontiguousarray(zs[:n])))
    blk_z[capacity] = float(n)
    out_s = torch.empty((world * capacity, sdim), dtype=torch.int8, device=dev)
    out_p = torch.empty((world * capacity, adim), dtype=torch.float32, device=dev)
    out_z = torch.empty((world * (capacity + 1),), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(out_s, blk_s, group=group)
    dist.all_gather_into_tensor(out_p, blk_p, group=group)
    dist.all_gather_into_tensor(out_z, blk_z, group=group)
    out_z = out_z.view(world, capacity + 1).cpu()
    counts = out_z[:, capacity].to(torch.int64).tolist()
    out_s = out_s.view(world, capacity, sdim).cpu().numpy()
    out_p = out_p.view(world, capacity, adim).cpu().numpy()
    zz = out_z.numpy()
    S = np.concatenate([out_s[r, :c] for r, c in enumerate(counts)], axis=0).reshape((-1,) + tuple(states.shape[1:]))
    P = np.concatenate([out_p[r, :c] for r, c in enumerate(counts)], axis=0)
    Z = np.concatenate([zz[r, :c] for r, c in enumerate(counts)], axis=0)
    return S, P, Z, kept


class SampleGatherer:
    """One all-gather per collection round with a fixed block per rank; what does not fit the block waits for the next round
    (finished games come in waves, the collective shape stays static).  `push` returns every rank's samples of this round."""

    def __init__(self, capacity, device=None, group=None):
        self.capacity, self.device, self.group = int(capacity), device, group
        self._backlog = None
        self.total = 0

    def pending(self):
        return 0 if self._backlog is None else len(self._backlog[2])

    def push(self, states, pis, zs):
        if self._backlog is not None:
            b = self._backlog
            states, pis, zs = np.concatenate([b[0], states]), np.concatenate([b[1], pis]), np.concatenate([b[2], zs])
            self._backlog = None
        S, P, Z, kept = all_gather_samples(states, pis, zs, self.capacity, device=self.device, group=self.group)
        if kept:
            self._backlog = (states[-kept:], pis[-kept:], zs[-kept:])
        self.total += len(Z)
        return S, P, Z
