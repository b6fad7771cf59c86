"""Constructor-only space classes for the gymnasium shim (see __init__.py)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = shape
        self.dtype = np.dtype(dtype) if dtype is not None else None

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        return True


class Discrete(Space):
    def __init__(self, n, start=0):
        super().__init__((), np.int64)
        self.n = int(n)
        self.start = int(start)

    def sample(self):
        return int(np.random.randint(self.start, self.start + self.n))


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(tuple(shape), dtype)
        self.low = low
        self.high = high


class MultiBinary(Space):
    def __init__(self, n):
        shape = (n,) if np.isscalar(n) else tuple(n)
        super().__init__(shape, np.int8)
        self.n = n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64):
        nvec = np.asarray(nvec)
        super().__init__(nvec.shape, dtype)
        self.nvec = nvec


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        super().__init__(None, None)
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()


class Tuple(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = tuple(spaces)
