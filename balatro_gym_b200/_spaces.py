"""Tiny stand-ins for gymnasium.spaces used ONLY when gymnasium is not installed
(this image has no gymnasium and no network).  Constructor-level compatibility:
Discrete(n).n, Box(low, high, shape, dtype), MultiBinary(n), Dict(mapping).spaces."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = np.dtype(dtype) if dtype is not None else None


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self, mask=None):
        if mask is not None:
            return int(np.random.choice(np.flatnonzero(mask)))
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(tuple(shape) if shape is not None else np.shape(low), dtype)
        self.low, self.high = low, high


class MultiBinary(Space):
    def __init__(self, n):
        super().__init__((n,), np.int8)
        self.n = n


class Dict(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()
