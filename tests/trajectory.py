"""Canonical byte serialisation of an env trajectory (must match tests/golden/make_golden.py:Trajectory)."""
import hashlib

import numpy as np


def parse_corpus(npz):
    off = npz['offsets']
    return [npz['moves'][off[i]:off[i + 1]].astype(np.int64) for i in range(len(off) - 1)]


class Trajectory:
    def __init__(self, keep=False):
        self.h = hashlib.sha1()
        self.keep = keep
        self.rows = []

    def add(self, legal, board, reward, done, to_play):
        legal = np.asarray(legal).astype(np.uint8)
        board = np.asarray(board).astype(np.int8).ravel()
        tail = np.array([int(reward), int(done), int(to_play)], dtype=np.int8)
        self.h.update(legal.tobytes())
        self.h.update(board.tobytes())
        self.h.update(tail.tobytes())
        if self.keep:
            self.rows.append((legal, board, tail))

    def hexdigest(self):
        return self.h.hexdigest()
