"""Output record of the self-play path: Transition(state, pi_prob, value) (reference: alpha_zero/core/replay.py:14-17).
When the reference package is importable (the training drivers run with it on the path) its own class is used so the
learner's replay sees exactly its type; otherwise an identical NamedTuple."""
from typing import NamedTuple, Optional

import numpy as np

try:  # pragma: no cover - depends on the caller's environment
    from alpha_zero.core.replay import Transition  # type: ignore
except Exception:  # noqa: BLE001

    class Transition(NamedTuple):
        state: Optional[np.ndarray]
        pi_prob: Optional[np.ndarray]
        value: Optional[float]


TRANSFORMATIONS = ['h_flip', 'v_flip', 'rotate90', 'rotate180', 'rotate270']  # utils/transformation.py:144-152, ids 1..5 in the C ABI


class DeviceReplay:
    """UniformReplay (core/replay.py:35-116) whose storage lives in the engine's HBM.

    Same surface — `add_game`, `add`, `sample(batch_size)`, `size`, `num_games_added`, `num_samples_added` — plus
    `ingest()` (finished self-play games go ring -> replay on the device, no host copy) and `sample(..., augment=True)` which
    fuses apply_random_transformation (utils/transformation.py:160) into the gather.  Host-side RNG calls are the reference's
    (`random_state.randint(low=0, high=size, size=batch)`; `random.random()` / `random.choice` for the augmentation), so seeded
    runs draw the same minibatches."""

    def __init__(self, engine, capacity, random_state):
        if capacity <= 0:
            raise ValueError(f'Expect capacity to be a positive integer, got {capacity}')
        self.engine = engine
        self.capacity = capacity
        self.random_state = random_state
        self.structure = Transition(state=None, pi_prob=None, value=None)
        engine.replay_create(capacity)

    def _info(self):
        return self.engine.replay_info()

    @property
    def size(self):
        return self._info()['size']

    @property
    def num_games_added(self):
        return self._info()['num_games_added']

    @property
    def num_samples_added(self):
        return self._info()['num_samples_added']

    def add_game(self, game_seq):
        if len(game_seq):
            self.engine.replay_add(np.stack([t.state for t in game_seq]), np.stack([np.asarray(t.pi_prob, dtype=np.float32) for t in game_seq]),
                                   np.array([t.value for t in game_seq], dtype=np.float32), n_games=1)

    def add(self, transition):
        self.engine.replay_add(transition.state[None], np.asarray(transition.pi_prob, dtype=np.float32)[None], np.array([transition.value], dtype=np.float32),
                               n_games=0)

    def ingest(self):
        """Move every finished self-play game of the engine into the replay (device to device). Returns (games, samples)."""
        return self.engine.replay_ingest()

    def sample(self, batch_size, augment=False, out=None):
        if self.size < batch_size:
            return None
        indices = self.random_state.randint(low=0, high=self.size, size=batch_size)
        transform = 0
        if augment:
            import random

            if random.random() > 0.5:
                transform = 1 + TRANSFORMATIONS.index(random.choice(TRANSFORMATIONS))
        res = self.engine.replay_sample(indices, transform, out=out)
        if out is not None:
            return True  # the batch sits in the caller's CUDA tensors; None stays reserved for "not enough samples yet"
        return type(self.structure)(state=res[0], pi_prob=res[1], value=res[2])
