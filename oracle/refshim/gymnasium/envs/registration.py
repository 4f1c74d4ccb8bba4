"""gymnasium.envs.registration stand-in: register / make / make_vec over callables."""
from __future__ import annotations

from dataclasses import dataclass, field

from .. import error

registry: dict = {}


@dataclass
class EnvSpec:
    id: str
    entry_point: object = None
    vector_entry_point: object = None
    kwargs: dict = field(default_factory=dict)
    max_episode_steps: int | None = None


def register(id, entry_point=None, vector_entry_point=None, kwargs=None, max_episode_steps=None, **_):
    registry[id] = EnvSpec(id, entry_point, vector_entry_point, dict(kwargs or {}), max_episode_steps)


def _spec(id):
    if id not in registry:
        raise error.NameNotFound(f"Environment `{id}` doesn't exist.")
    return registry[id]


def make(id, max_episode_steps=None, disable_env_checker=None, **kwargs):
    spec = _spec(id)
    if spec.entry_point is None:
        raise error.Error(f"{id} registered but entry_point is not specified")
    kw = dict(spec.kwargs)
    kw.update(kwargs)
    env = spec.entry_point(**kw)
    return env


def make_vec(id, num_envs=1, vectorization_mode=None, vector_kwargs=None, wrappers=None, **kwargs):
    spec = _spec(id)
    if spec.vector_entry_point is None:
        raise error.Error(f"{id} registered but vector_entry_point is not specified")
    kw = dict(spec.kwargs)
    kw.update(kwargs)
    kw.update(vector_kwargs or {})
    # gymnasium passes num_envs through to a custom vector entry point
    return spec.vector_entry_point(num_envs=num_envs, **kw)
