"""gymnasium.spaces stand-in: Space, Box, Discrete with gymnasium 1.x sampling semantics."""
from __future__ import annotations

import numpy as np

from ..utils import seeding


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = None
        self._np_random_seed = None
        if seed is not None:
            if isinstance(seed, np.random.Generator):
                self._np_random = seed
            else:
                self.seed(seed)

    @property
    def np_random(self):
        if self._np_random is None:
            self.seed()
        return self._np_random

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._np_random, self._np_random_seed = seeding.np_random(seed)
        return self._np_random_seed

    def sample(self, mask=None):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)


def _broadcast(value, dtype, shape):
    if np.isscalar(value):
        return np.full(shape, value, dtype=dtype)
    return np.asarray(value).astype(dtype)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.asarray(low).shape if not np.isscalar(low) else np.asarray(high).shape
        shape = tuple(int(d) for d in shape)
        self.low = _broadcast(low, dtype, shape)
        self.high = _broadcast(high, dtype, shape)
        assert self.low.shape == shape and self.high.shape == shape
        self.bounded_below = -np.inf < self.low
        self.bounded_above = np.inf > self.high
        super().__init__(shape, dtype, seed)

    def sample(self, mask=None):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        sample = np.empty(self.shape)
        unbounded = ~self.bounded_below & ~self.bounded_above
        upp_bounded = ~self.bounded_below & self.bounded_above
        low_bounded = self.bounded_below & ~self.bounded_above
        bounded = self.bounded_below & self.bounded_above
        sample[unbounded] = self.np_random.normal(size=unbounded[unbounded].shape)
        sample[low_bounded] = self.np_random.exponential(size=low_bounded[low_bounded].shape) + self.low[low_bounded]
        sample[upp_bounded] = -self.np_random.exponential(size=upp_bounded[upp_bounded].shape) + high[upp_bounded]
        sample[bounded] = self.np_random.uniform(low=self.low[bounded], high=high[bounded], size=bounded[bounded].shape)
        if self.dtype.kind in ["i", "u", "b"]:
            sample = np.floor(sample)
        return sample.astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                and np.allclose(self.low, other.low) and np.allclose(self.high, other.high))


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = np.int64(n)
        self.start = np.int64(start)
        super().__init__((), np.int64, seed)

    def sample(self, mask=None):
        return self.start + self.np_random.integers(self.n)

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n
