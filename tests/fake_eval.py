"""Deterministic stand-in evaluator shared by the golden generator, the oracle tests and the GPU parity tests.

It maps an observation to (prior, value) through a CRC of the observation bytes, so the reference
search (mcts_v2.py:301,485), the oracle restatement and the CUDA tree all see the *same* (P, v) stream
without a neural net in the loop.  The return types follow the reference's `eval_position`
contract (pipeline.py:91-123): float32 1-D priors, Python-float values holding float32 numbers.
"""
import zlib

import numpy as np


def make_fake_eval(num_actions: int, sharpness: float = 2.0):
    def _one(obs: np.ndarray):
        assert obs.dtype == np.int8, obs.dtype
        seed = zlib.crc32(np.ascontiguousarray(obs).tobytes())
        rng = np.random.RandomState(seed)
        logits = (rng.standard_normal(num_actions) * sharpness).astype(np.float32)
        logits -= logits.max()
        p = np.exp(logits).astype(np.float32)
        p = (p / p.sum(dtype=np.float32)).astype(np.float32)
        v = float(np.float32(rng.uniform(-1.0, 1.0)))
        return p, v

    def eval_func(obs: np.ndarray, batched: bool = False):
        if not batched:
            return _one(obs)
        out = [_one(o) for o in obs]
        return [o[0] for o in out], [o[1] for o in out]

    return eval_func
