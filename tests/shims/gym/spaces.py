"""Stub of gym.spaces: only the attributes the reference touches."""


class Box:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Discrete:
    def __init__(self, n):
        self.n = int(n)
