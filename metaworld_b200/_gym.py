"""Gymnasium interop.  When ``gymnasium`` is importable the vector env subclasses its ``VectorEnv`` and uses
its ``Box``; otherwise a minimal stand-in with the same surface (low/high/shape/dtype/sample/seed/contains)
is used so the engine has no hard dependency on it (the reference requires gymnasium>=1.1, pyproject.toml:27)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _g
    from gymnasium.spaces import Box as _GBox
    from gymnasium.vector.utils import batch_space as _gbatch

    HAVE_GYMNASIUM = True
    VectorEnvBase = _g.vector.VectorEnv

    def Box(low, high, dtype=np.float32, seed=None):
        return _GBox(low, high, dtype=dtype, seed=seed)

    def batch_space(space, n, seed=None):
        s = _gbatch(space, n)
        if seed is not None:
            s.seed(seed)
        return s

except Exception:  # gymnasium not installed
    HAVE_GYMNASIUM = False

    class VectorEnvBase:  # noqa: D401 - minimal protocol holder
        """Stand-in for gymnasium.vector.VectorEnv."""

        closed = False

        @property
        def unwrapped(self):
            return self

    class _Box:
        def __init__(self, low, high, dtype=np.float32, seed=None):
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return seed

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def Box(low, high, dtype=np.float32, seed=None):
        return _Box(low, high, dtype=dtype, seed=seed)

    def batch_space(space, n, seed=None):
        return _Box(np.repeat(space.low[None], n, 0), np.repeat(space.high[None], n, 0), dtype=space.dtype, seed=seed)
