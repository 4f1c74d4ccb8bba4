"""gymnasium.vector stand-in: SyncVectorEnv with AutoresetMode (NEXT_STEP / SAME_STEP / DISABLED) and the
dict-of-arrays info aggregation (`_add_info`: per-key array + `_key` mask) of gymnasium 1.1."""
from __future__ import annotations

from copy import deepcopy
from enum import Enum

import numpy as np

from ..spaces import Box


class AutoresetMode(Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


def batch_space(space, n):
    assert isinstance(space, Box)
    rep = (n,) + (1,) * space.low.ndim
    return Box(low=np.tile(space.low, rep), high=np.tile(space.high, rep), dtype=space.dtype,
               seed=deepcopy(space.np_random))


class VectorEnv:
    metadata: dict = {}
    spec = None
    render_mode = None
    closed = False
    num_envs: int
    _np_random = None

    def close(self, **kw):
        if not self.closed:
            self.close_extras(**kw)
            self.closed = True

    def close_extras(self, **kw):
        pass

    @property
    def unwrapped(self):
        return self

    def _add_info(self, vector_infos, env_info, env_num):
        for key, value in env_info.items():
            if isinstance(value, dict):
                array = self._add_info(vector_infos.get(key, {}), value, env_num)
            else:
                if key not in vector_infos:
                    if type(value) in [int, float, bool] or issubclass(type(value), np.number):
                        array = np.zeros(self.num_envs, dtype=type(value))
                    elif isinstance(value, np.ndarray):
                        array = np.zeros((self.num_envs, *value.shape), dtype=value.dtype)
                    else:
                        array = np.full(self.num_envs, fill_value=None, dtype=object)
                else:
                    array = vector_infos[key]
                array[env_num] = value
            array_mask = vector_infos.get(f"_{key}", np.zeros(self.num_envs, dtype=np.bool_))
            array_mask[env_num] = True
            vector_infos[key], vector_infos[f"_{key}"] = array, array_mask
        return vector_infos


class SyncVectorEnv(VectorEnv):
    def __init__(self, env_fns, copy=True, observation_mode="same", autoreset_mode=AutoresetMode.NEXT_STEP):
        self.copy = copy
        self.env_fns = env_fns
        self.autoreset_mode = autoreset_mode if isinstance(autoreset_mode, AutoresetMode) else AutoresetMode(autoreset_mode)
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.metadata = dict(self.envs[0].metadata)
        self.metadata["autoreset_mode"] = self.autoreset_mode
        self.render_mode = self.envs[0].render_mode
        self.single_action_space = self.envs[0].action_space
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.single_observation_space = self.envs[0].observation_space
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        self._env_obs = [None] * self.num_envs
        self._observations = np.zeros(self.observation_space.shape, dtype=self.observation_space.dtype)
        self._rewards = np.zeros((self.num_envs,), dtype=np.float64)
        self._terminations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._truncations = np.zeros((self.num_envs,), dtype=np.bool_)
        self._autoreset_envs = np.zeros((self.num_envs,), dtype=np.bool_)

    @property
    def np_random_seed(self):
        return self.get_attr("np_random_seed")

    @property
    def np_random(self):
        return self.get_attr("np_random")

    def _stack(self):
        for i, o in enumerate(self._env_obs):
            self._observations[i] = o
        return deepcopy(self._observations) if self.copy else self._observations

    def reset(self, *, seed=None, options=None):
        if seed is None:
            seed = [None] * self.num_envs
        elif isinstance(seed, int):
            seed = [seed + i for i in range(self.num_envs)]
        assert len(seed) == self.num_envs
        if options is not None and "reset_mask" in options:
            mask = options.pop("reset_mask")
            self._terminations[mask] = False
            self._truncations[mask] = False
            self._autoreset_envs[mask] = False
            infos = {}
            for i, (env, s, m) in enumerate(zip(self.envs, seed, mask)):
                if m:
                    self._env_obs[i], env_info = env.reset(seed=s, options=options)
                    infos = self._add_info(infos, env_info, i)
        else:
            self._terminations[:] = False
            self._truncations[:] = False
            self._autoreset_envs[:] = False
            infos = {}
            for i, (env, s) in enumerate(zip(self.envs, seed)):
                self._env_obs[i], env_info = env.reset(seed=s, options=options)
                infos = self._add_info(infos, env_info, i)
        return self._stack(), infos

    def step(self, actions):
        infos = {}
        for i, action in enumerate(actions):
            if self.autoreset_mode == AutoresetMode.NEXT_STEP:
                if self._autoreset_envs[i]:
                    self._env_obs[i], env_info = self.envs[i].reset()
                    self._rewards[i], self._terminations[i], self._truncations[i] = 0.0, False, False
                else:
                    (self._env_obs[i], self._rewards[i], self._terminations[i], self._truncations[i],
                     env_info) = self.envs[i].step(action)
            elif self.autoreset_mode == AutoresetMode.DISABLED:
                assert not self._autoreset_envs[i], f"{self._autoreset_envs=}"
                (self._env_obs[i], self._rewards[i], self._terminations[i], self._truncations[i],
                 env_info) = self.envs[i].step(action)
            elif self.autoreset_mode == AutoresetMode.SAME_STEP:
                (self._env_obs[i], self._rewards[i], self._terminations[i], self._truncations[i],
                 env_info) = self.envs[i].step(action)
                if self._terminations[i] or self._truncations[i]:
                    infos = self._add_info(infos, {"final_obs": self._env_obs[i], "final_info": env_info}, i)
                    self._env_obs[i], env_info = self.envs[i].reset()
            else:
                raise ValueError(self.autoreset_mode)
            infos = self._add_info(infos, env_info, i)
        self._autoreset_envs = np.logical_or(self._terminations, self._truncations)
        return (self._stack(), np.copy(self._rewards), np.copy(self._terminations), np.copy(self._truncations), infos)

    def call(self, name, *args, **kwargs):
        results = []
        for env in self.envs:
            fn = env.get_wrapper_attr(name)
            results.append(fn(*args, **kwargs) if callable(fn) else fn)
        return tuple(results)

    def get_attr(self, name):
        return self.call(name)

    def set_attr(self, name, values):
        if not isinstance(values, (list, tuple)):
            values = [values for _ in range(self.num_envs)]
        assert len(values) == self.num_envs
        for env, v in zip(self.envs, values):
            env.set_wrapper_attr(name, v)

    def close_extras(self, **kw):
        for env in getattr(self, "envs", []):
            env.close()


class AsyncVectorEnv(SyncVectorEnv):
    """The shim runs 'async' in-process (no subprocesses); results are identical to sync by construction."""
