"""metaworld_b200: Blackwell-native batched Meta-World step engine.

Public surface mirrors the reference package (``metaworld/__init__.py``): ``make_mt_envs`` / ``make_ml_envs``
and, when gymnasium is installed, the ``Meta-World/MT1 | MT10 | MT25 | MT50 | ML1-* | ML10-* | ML25-* | ML45-*``
ids registered for ``gym.make_vec`` (namespace ``Meta-World-B200/`` so both packages can coexist)."""
from __future__ import annotations

__version__ = "0.1.0"

from .benchmarks import ALL_V3, ML1, ML10, ML25, ML45, MT1, MT10, MT25, MT50, Benchmark, Task, make_benchmark  # noqa: F401
from . import evaluation  # noqa: F401  (evaluation / metalearning_evaluation, metaworld/evaluation.py)


def make_mt_envs(*a, **k):
    from .vector_env import make_mt_envs as f
    return f(*a, **k)


def make_ml_envs(*a, **k):
    from .vector_env import make_ml_envs as f
    return f(*a, **k)


def register_mw_envs():
    """gymnasium registration (metaworld/__init__.py:607-820) when gymnasium is available."""
    from . import _gym
    if not _gym.HAVE_GYMNASIUM:
        return False
    from gymnasium.envs.registration import register

    def vec(name, ml_split=None):
        def entry(seed=None, num_envs=None, **kw):
            if ml_split is None:
                return make_mt_envs(kw.pop("env_name", name), seed=seed, num_envs=num_envs, **kw)
            return make_ml_envs(kw.pop("env_name", name), seed=seed, split=ml_split, num_envs=num_envs, **kw)
        return entry

    for n in ("MT1", "MT10", "MT25", "MT50"):
        register(id=f"Meta-World-B200/{n}", vector_entry_point=vec(n), kwargs={})
    for n in ("ML1", "ML10", "ML25", "ML45"):
        for split in ("train", "test"):
            register(id=f"Meta-World-B200/{n}-{split}", vector_entry_point=vec(n, split), kwargs={})
    return True


try:
    register_mw_envs()
except Exception:  # registration is best effort
    pass
