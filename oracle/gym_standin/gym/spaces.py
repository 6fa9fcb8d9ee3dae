import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]


class Box(Space):
    """Bounded box; `sample()` is per-dimension uniform(low, high) from the space's
    own RandomState and `contains` is shape match + low <= x <= high (gym 0.17)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            assert low.shape == high.shape
            shape = low.shape
            self.low = low.astype(dtype)
            self.high = high.astype(dtype)
        else:
            shape = tuple(shape)
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
        super(Box, self).__init__(shape, dtype)

    def sample(self):
        return self.np_random.uniform(low=self.low, high=self.high,
                                      size=self.shape).astype(self.dtype)

    def contains(self, x):
        if isinstance(x, list):
            x = np.array(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        super(Discrete, self).__init__((), np.int64)

    def sample(self):
        return self.np_random.randint(self.n)

    def contains(self, x):
        return 0 <= int(x) < self.n
