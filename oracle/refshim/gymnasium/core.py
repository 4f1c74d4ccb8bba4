"""gymnasium.core stand-in: Env / Wrapper / ObservationWrapper (gymnasium 1.x: wrappers do NOT forward arbitrary
attributes; `np_random` of a wrapper is the wrapped env's)."""
from __future__ import annotations

from .utils import seeding


class Env:
    metadata: dict = {"render_modes": []}
    render_mode = None
    spec = None
    action_space = None
    observation_space = None
    _np_random = None
    _np_random_seed = None

    def step(self, action):
        raise NotImplementedError

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = seeding.np_random(seed)
        return None

    def render(self):
        return None

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self

    @property
    def np_random_seed(self):
        if self._np_random_seed is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random_seed

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = seeding.np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value
        self._np_random_seed = -1

    def has_wrapper_attr(self, name):
        return hasattr(self, name)

    def get_wrapper_attr(self, name):
        return getattr(self, name)

    def set_wrapper_attr(self, name, value, *, force=True):
        if force or hasattr(self, name):
            setattr(self, name, value)
            return True
        return False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None
        self._metadata = None

    def step(self, action):
        return self.env.step(action)

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def np_random(self):
        return self.env.np_random

    @np_random.setter
    def np_random(self, value):
        self.env.np_random = value

    @property
    def np_random_seed(self):
        return self.env.np_random_seed

    @property
    def action_space(self):
        return self.env.action_space if self._action_space is None else self._action_space

    @action_space.setter
    def action_space(self, s):
        self._action_space = s

    @property
    def observation_space(self):
        return self.env.observation_space if self._observation_space is None else self._observation_space

    @observation_space.setter
    def observation_space(self, s):
        self._observation_space = s

    @property
    def metadata(self):
        return self.env.metadata if self._metadata is None else self._metadata

    @metadata.setter
    def metadata(self, v):
        self._metadata = v

    @property
    def render_mode(self):
        return self.env.render_mode

    @property
    def spec(self):
        return self.env.spec

    def has_wrapper_attr(self, name):
        return hasattr(self, name) or self.env.has_wrapper_attr(name)

    def get_wrapper_attr(self, name):
        if hasattr(self, name):
            return getattr(self, name)
        try:
            return self.env.get_wrapper_attr(name)
        except AttributeError as e:
            raise AttributeError(f"wrapper {type(self).__name__} has no attribute {name!r}") from e

    def set_wrapper_attr(self, name, value, *, force=True):
        if hasattr(self, name):
            setattr(self, name, value)
            return True
        done = self.env.set_wrapper_attr(name, value, force=False)
        if done:
            return True
        if force:
            setattr(self, name, value)
            return True
        return False


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, obs):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return obs, self.reward(reward), terminated, truncated, info


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))
