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
