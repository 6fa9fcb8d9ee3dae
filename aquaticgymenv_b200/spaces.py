"""Minimal ``gym.spaces`` surface used at the reference's call sites (``Box``: low/high/shape/
sample/contains, ``Discrete``: n/sample/contains -- aqua.py:30-31,43,46-53; dqn.py:99-100;
utils.py:19-21; test_random.py:20).  ``gym`` itself is a third-party package that is not part of
the reference tree; if it is installed, the real spaces are used instead (see gym_aqua/__init__)."""
import numpy as np


class Space(object):
    def __init__(self, shape, dtype, seed=None):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.np_random = np.random.RandomState(seed)

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            shape = low.shape
        else:
            shape = tuple(shape)
            low = np.full(shape, low, dtype=dtype)
            high = np.full(shape, high, dtype=dtype)
        super().__init__(shape, dtype)
        self.low, self.high = low, high

    def sample(self):
        return self.np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        try:
            return 0 <= int(x) < self.n and int(x) == x
        except (TypeError, ValueError):
            return False

    def __repr__(self):
        return "Discrete(%d)" % self.n
