"""gymnasium.wrappers stand-ins on the reference's step path: TimeLimit, RecordEpisodeStatistics
(+ NormalizeReward / NormalizeObservation, used only behind non-default kwargs)."""
from __future__ import annotations

import time
from collections import deque

import numpy as np

from ..core import ObservationWrapper, Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        assert isinstance(max_episode_steps, (int, np.integer)) and max_episode_steps > 0
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    def reset(self, *, seed=None, options=None):
        self._elapsed_steps = 0
        return super().reset(seed=seed, options=options)


class RecordEpisodeStatistics(Wrapper):
    def __init__(self, env, buffer_length=100, stats_key="episode"):
        super().__init__(env)
        self._stats_key = stats_key
        self.episode_count = 0
        self.episode_start_time = -1.0
        self.episode_returns = 0.0
        self.episode_lengths = 0
        self.time_queue = deque(maxlen=buffer_length)
        self.return_queue = deque(maxlen=buffer_length)
        self.length_queue = deque(maxlen=buffer_length)

    def step(self, action):
        obs, reward, terminated, truncated, info = super().step(action)
        self.episode_returns += reward
        self.episode_lengths += 1
        if terminated or truncated:
            assert self._stats_key not in info
            elapsed = round(time.perf_counter() - self.episode_start_time, 6)
            info[self._stats_key] = {"r": self.episode_returns, "l": self.episode_lengths, "t": elapsed}
            self.time_queue.append(elapsed)
            self.return_queue.append(self.episode_returns)
            self.length_queue.append(self.episode_lengths)
            self.episode_count += 1
            self.episode_start_time = time.perf_counter()
        return obs, reward, terminated, truncated, info

    def reset(self, *, seed=None, options=None):
        obs, info = super().reset(seed=seed, options=options)
        self.episode_start_time = time.perf_counter()
        self.episode_returns = 0.0
        self.episode_lengths = 0
        return obs, info


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=(), dtype=np.float64):
        self.mean = np.zeros(shape, dtype=dtype)
        self.var = np.ones(shape, dtype=dtype)
        self.count = epsilon

    def update(self, x):
        bm, bv, bc = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
        delta = bm - self.mean
        tot = self.count + bc
        self.mean = self.mean + delta * bc / tot
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.var = m2 / tot
        self.count = tot


class NormalizeReward(Wrapper):
    def __init__(self, env, gamma=0.99, epsilon=1e-8):
        super().__init__(env)
        self.return_rms = RunningMeanStd(shape=())
        self.discounted_reward = np.array([0.0])
        self.gamma, self.epsilon, self._update_running_mean = gamma, epsilon, True

    def step(self, action):
        obs, reward, terminated, truncated, info = super().step(action)
        self.discounted_reward = self.discounted_reward * self.gamma * (1 - terminated) + float(reward)
        if self._update_running_mean:
            self.return_rms.update(self.discounted_reward)
        return obs, reward / np.sqrt(self.return_rms.var + self.epsilon), terminated, truncated, info


class NormalizeObservation(ObservationWrapper):
    def __init__(self, env, epsilon=1e-8):
        super().__init__(env)
        from ..spaces import Box
        shape = env.observation_space.shape
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=shape, dtype=np.float32)
        self.obs_rms = RunningMeanStd(shape=shape, dtype=np.float32)
        self.epsilon, self._update_running_mean = epsilon, True

    def observation(self, observation):
        if self._update_running_mean:
            self.obs_rms.update(np.array([observation]))
        return np.float32((observation - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon))
