"""Per-process pools of engine slots for the single-env façades.

A `GoEnv()` / `GomokuEnv()` object is a handle on one slot of an Engine created lazily in the process
that first touches it (the training drivers pickle env objects into `spawn`ed actors, so nothing
CUDA-related may exist at construction / unpickling time).
"""
import os

_POOLS = {}
_TEST_BINDING = None  # tests may inject the host-emulation binding; the product path never sets this


class SlotPool:
    def __init__(self, key):
        from ..engine import Engine

        game, n, komi, max_steps, num_to_win, num_stack = key
        slots = int(os.environ.get('AZ_POOL_SLOTS', '32'))
        self.max_sims = int(os.environ.get('AZ_POOL_MAX_SIMULATIONS', '1600' if n <= 13 else '1000'))
        self.max_parallel = int(os.environ.get('AZ_POOL_MAX_PARALLEL', '16'))
        self.engine = Engine(game, n, num_games=slots, max_simulations=self.max_sims, max_parallel=self.max_parallel, komi=komi,
                             max_steps=max_steps, num_to_win=num_to_win, num_stack=num_stack, net=None,
                             device=int(os.environ.get('AZ_DEVICE', '0')), binding=_TEST_BINDING)
        self.free = list(range(slots - 1, -1, -1))

    def acquire(self):
        if not self.free:
            raise RuntimeError('engine slot pool exhausted: raise AZ_POOL_SLOTS (envs are released when garbage collected)')
        return self.free.pop()

    def release(self, slot):
        self.free.append(slot)


def get_pool(key):
    if key not in _POOLS:
        _POOLS[key] = SlotPool(key)
    return _POOLS[key]


def reset_pools():
    for p in _POOLS.values():
        p.engine.close()
    _POOLS.clear()
