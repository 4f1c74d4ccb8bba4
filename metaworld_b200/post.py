"""Vectorised equivalents of the two per-sub-env wrappers the reference stacks between the one-hot wrapper and the
episode-statistics wrapper (metaworld/__init__.py:437-444):

* `RNNBasedMetaRLWrapper` (metaworld/wrappers.py:35-88): obs <- [obs, action, reward (/10), done]; after a reset the three
  extra blocks are zero;
* `NormalizeRewardsExponential` (wrappers.py:233-258): per-env exponential running mean / variance of the reward; the
  reference updates the estimate twice per step (once in `step`, once more inside `_apply_normalize_reward`), which is
  reproduced here because it changes the numbers.

* `gymnasium.wrappers.NormalizeReward` (reward_normalization_method="gymnasium", metaworld/__init__.py:441-442): reward
  divided by the running standard deviation of the discounted return (gamma 0.99; the return is zeroed on termination
  only, and never on reset);
* `gymnasium.wrappers.NormalizeObservation` (normalize_observations=True, metaworld/__init__.py:445-446): per-sub-env
  running mean / variance (Welford merge of one-sample batches, initial count 1e-4) over every observation the wrapper
  sees, i.e. step observations AND reset observations; output float32.

Pure numpy on [N, ...] arrays; `MetaWorldVecEnv` applies it to what the engine returns.  Because the reference puts
`RecordEpisodeStatistics` outside the reward normalisation, the episodic return it reports is the sum of NORMALISED
rewards; `ep_return` below tracks that."""
from __future__ import annotations

import numpy as np


class RunningMeanStdBatch:
    """N independent copies of gymnasium's RunningMeanStd (wrappers/utils.py), each fed one sample per update: the
    parallel-variance merge with batch mean = x, batch variance = 0, batch count = 1.  Arithmetic runs in `dtype`, with the
    count converted at each use, like a Python-float count combined with float32 arrays in the per-env wrapper."""

    def __init__(self, n, shape, dtype, epsilon=1e-4):
        self.mean = np.zeros((n,) + tuple(shape), dtype=dtype)
        self.var = np.ones((n,) + tuple(shape), dtype=dtype)
        self.count = np.full(n, epsilon, dtype=np.float64)
        self._bc = (n,) + (1,) * len(shape)

    def update(self, x, mask=None):
        dt = self.mean.dtype
        count = self.count.reshape(self._bc)
        tot = count + 1.0
        c, tt = count.astype(dt), tot.astype(dt)
        delta = np.asarray(x, dtype=dt) - self.mean
        mean = self.mean + delta / tt
        m2 = self.var * c + np.square(delta) * c / tt
        var = m2 / tt
        if mask is None:
            self.mean, self.var, self.count = mean, var, self.count + 1.0
        else:
            m = np.asarray(mask, dtype=bool)
            mb = m.reshape(self._bc)
            self.mean, self.var = np.where(mb, mean, self.mean), np.where(mb, var, self.var)
            self.count = np.where(m, self.count + 1.0, self.count)


class StepPost:
    GAMMA, EPS = 0.99, 1e-8          # gymnasium's defaults; the reference passes none

    def __init__(self, num_envs, recurrent_info_in_obs=False, normalize_reward_in_recurrent_info=True,
                 reward_normalization_method=None, reward_alpha=0.001, normalize_observations=False):
        if reward_normalization_method not in (None, "exponential", "gymnasium"):
            raise ValueError(f"unknown reward_normalization_method {reward_normalization_method!r}")
        self.n = num_envs
        self.gym_reward = reward_normalization_method == "gymnasium"
        self.norm_obs = bool(normalize_observations)
        self.ret = RunningMeanStdBatch(num_envs, (), np.float64)      # NormalizeReward.return_rms of every sub-env
        self.disc = np.zeros(num_envs)                                 # NormalizeReward.discounted_reward
        self.obs_rms = None                                            # NormalizeObservation.obs_rms, shaped at the first observation
        self.recurrent = bool(recurrent_info_in_obs)
        self.norm_in_obs = bool(normalize_reward_in_recurrent_info)
        self.exponential = reward_normalization_method == "exponential"
        self.alpha = float(reward_alpha)
        self.mean = np.zeros(num_envs)
        self.var = np.ones(num_envs)
        self.ep_return = np.zeros(num_envs)
        self.extra = 6 if self.recurrent else 0

    @property
    def active(self):
        return self.recurrent or self.exponential or self.gym_reward or self.norm_obs

    def _normalize_obs(self, obs, mask=None):
        """NormalizeObservation.observation on the rows selected by `mask` (all when None): update, then normalise."""
        if self.obs_rms is None:
            self.obs_rms = RunningMeanStdBatch(self.n, obs.shape[1:], obs.dtype)
        self.obs_rms.update(obs, mask)
        return np.float32((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.EPS))

    def on_reset(self, obs, mask=None):
        """obs [N, D] -> [N, D + extra]; the reward statistics are NOT reset (the wrapper object lives across episodes)."""
        if mask is None:
            self.ep_return[:] = 0
        else:
            self.ep_return[mask] = 0
        if self.recurrent:
            obs = np.concatenate([obs, np.zeros((len(obs), 6), dtype=obs.dtype)], axis=1)
        if self.norm_obs:
            obs = self._normalize_obs(obs, mask)
        return obs

    def _update(self, r):
        self.mean = (1 - self.alpha) * self.mean + self.alpha * r
        self.var = (1 - self.alpha) * self.var + self.alpha * np.square(r - self.mean)

    def on_step(self, obs, actions, reward, terminated, truncated, final_obs=None):
        """Returns (obs_out, reward_out, final_obs_out, episode_return_of_finished_envs).
        `obs` holds the post-autoreset observation for finished envs (SAME_STEP) and `final_obs` their terminal one."""
        done = np.logical_or(terminated, truncated)
        obs_out, final_out = obs, final_obs
        if self.recurrent:
            r_obs = reward / 10.0 if self.norm_in_obs else reward
            ext = np.concatenate([np.asarray(actions, dtype=obs.dtype).reshape(self.n, 4), r_obs[:, None].astype(obs.dtype),
                                  done[:, None].astype(obs.dtype)], axis=1)
            if final_obs is not None:
                final_out = np.concatenate([final_obs, ext], axis=1)
            ext = np.where(done[:, None], 0, ext).astype(obs.dtype)          # a freshly reset env reports zeros
            obs_out = np.concatenate([obs, ext], axis=1)
        reward_out = reward
        if self.exponential:
            self._update(reward)
            self._update(reward)
            reward_out = reward / (np.sqrt(self.var) + 1e-8)
        elif self.gym_reward:
            self.disc = self.disc * self.GAMMA * (1 - terminated) + reward
            self.ret.update(self.disc)
            reward_out = reward / np.sqrt(self.ret.var + self.EPS)
        if self.norm_obs:
            # the wrapper sees the step observation of every env (the terminal one for a finished env), then - SAME_STEP -
            # the reset observation of the finished ones: two updates for those, in that order
            if final_out is not None and done.any():
                stepped = np.where(done[:, None], final_out, obs_out)
                normed = self._normalize_obs(stepped)
                final_out = normed
                obs_out = np.where(done[:, None], self._normalize_obs(obs_out, done), normed)
            else:
                obs_out = self._normalize_obs(obs_out)
        self.ep_return += reward_out
        finished = np.where(done, self.ep_return, 0.0)
        self.ep_return[done] = 0
        return obs_out, reward_out, final_out, finished


class StepPostTorch:
    """The same two wrappers on CUDA tensors, for `MetaWorldVecEnv.step_torch` (so RL code that keeps everything on the GPU
    gets the recurrent observation and the normalised reward without a host round trip).  Plain torch elementwise ops on
    [N, ...] tensors: plumbing around the engine's outputs, not a hot path."""

    def __init__(self, torch, device, num_envs, obs_dim, recurrent_info_in_obs=False, normalize_reward_in_recurrent_info=True,
                 reward_normalization_method=None, reward_alpha=0.001, normalize_observations=False):
        self.t = torch
        self.gym_reward = reward_normalization_method == "gymnasium"
        self.norm_obs = bool(normalize_observations)
        D = obs_dim + (6 if recurrent_info_in_obs else 0)
        f64 = dict(device=device, dtype=torch.float64)
        self.ret_mean, self.ret_var = torch.zeros(num_envs, **f64), torch.ones(num_envs, **f64)
        self.ret_count = torch.full((num_envs,), 1e-4, **f64)
        self.disc = torch.zeros(num_envs, **f64)
        self.obs_mean, self.obs_var = torch.zeros(num_envs, D, **f64), torch.ones(num_envs, D, **f64)
        self.obs_count = torch.full((num_envs, 1), 1e-4, **f64)
        self.recurrent = bool(recurrent_info_in_obs)
        self.norm_in_obs = bool(normalize_reward_in_recurrent_info)
        self.exponential = reward_normalization_method == "exponential"
        self.alpha = float(reward_alpha)
        self.mean = torch.zeros(num_envs, device=device, dtype=torch.float64)
        self.var = torch.ones(num_envs, device=device, dtype=torch.float64)
        self.ep_return = torch.zeros(num_envs, device=device, dtype=torch.float64)
        self.obs_dim = obs_dim
        self.out = torch.zeros(num_envs, obs_dim + 6, device=device) if self.recurrent else None
        self.final_out = torch.zeros(num_envs, obs_dim + 6, device=device) if self.recurrent else None

    def load_host_state(self, post: StepPost):
        """Continue from the numpy-path statistics (a run may mix `step` and `step_torch`)."""
        t = self.t
        self.mean.copy_(t.from_numpy(post.mean)); self.var.copy_(t.from_numpy(post.var)); self.ep_return.copy_(t.from_numpy(post.ep_return))
        self.ret_mean.copy_(t.from_numpy(post.ret.mean)); self.ret_var.copy_(t.from_numpy(post.ret.var))
        self.ret_count.copy_(t.from_numpy(post.ret.count)); self.disc.copy_(t.from_numpy(post.disc))
        if post.obs_rms is not None:
            self.obs_mean.copy_(t.from_numpy(post.obs_rms.mean.astype(np.float64))); self.obs_var.copy_(t.from_numpy(post.obs_rms.var.astype(np.float64)))
            self.obs_count.copy_(t.from_numpy(post.obs_rms.count)[:, None])

    @staticmethod
    def _merge(mean, var, count, x):
        """one-sample parallel-variance merge (RunningMeanStdBatch.update); count broadcasts over the trailing axes"""
        tot = count + 1.0
        delta = x - mean
        return mean + delta / tot, (var * count + delta * delta * count / tot) / tot, tot

    def _normalize_obs(self, obs, mask=None):
        t = self.t
        x = obs.double()
        mean, var, count = self._merge(self.obs_mean, self.obs_var, self.obs_count, x)
        if mask is None:
            self.obs_mean, self.obs_var, self.obs_count = mean, var, count
        else:
            m = mask[:, None]
            self.obs_mean, self.obs_var, self.obs_count = t.where(m, mean, self.obs_mean), t.where(m, var, self.obs_var), t.where(m, count, self.obs_count)
        return ((x - self.obs_mean) / (self.obs_var + StepPost.EPS).sqrt()).float()

    def on_reset(self, obs):
        self.ep_return.zero_()
        if self.recurrent:
            self.out.zero_(); self.out[:, : self.obs_dim] = obs
            obs = self.out
        if self.norm_obs:
            obs = self._normalize_obs(obs)
        return obs

    def on_step(self, obs, actions, reward, terminated, truncated, final_obs):
        """-> (obs_out, reward_out [float64], final_obs_out, episode_return_of_finished_envs)"""
        t = self.t
        done = (terminated | truncated).bool()
        r64 = reward.double()
        obs_out, final_out = obs, final_obs
        if self.recurrent:
            r_obs = (r64 / 10.0 if self.norm_in_obs else r64).float()
            ext = t.cat([actions, r_obs[:, None], done[:, None].float()], dim=1)
            self.final_out[:, : self.obs_dim] = final_obs; self.final_out[:, self.obs_dim:] = ext
            self.out[:, : self.obs_dim] = obs; self.out[:, self.obs_dim:] = t.where(done[:, None], t.zeros_like(ext), ext)
            obs_out, final_out = self.out, self.final_out
        reward_out = r64
        if self.exponential:
            for _ in range(2):          # the reference updates the estimate twice per step (wrappers.py:250-258)
                self.mean = (1 - self.alpha) * self.mean + self.alpha * r64
                self.var = (1 - self.alpha) * self.var + self.alpha * (r64 - self.mean) ** 2
            reward_out = r64 / (self.var.sqrt() + 1e-8)
        elif self.gym_reward:
            self.disc = self.disc * StepPost.GAMMA * (1.0 - terminated.double()) + r64
            self.ret_mean, self.ret_var, self.ret_count = self._merge(self.ret_mean, self.ret_var, self.ret_count, self.disc)
            reward_out = r64 / (self.ret_var + StepPost.EPS).sqrt()
        if self.norm_obs:       # every env's step observation first, then the reset observation of the finished ones (no sync: masked)
            normed = self._normalize_obs(t.where(done[:, None], final_out, obs_out))
            obs_out = t.where(done[:, None], self._normalize_obs(obs_out, done), normed)
            final_out = normed
        self.ep_return += reward_out
        finished = t.where(done, self.ep_return, t.zeros_like(self.ep_return))
        self.ep_return = t.where(done, t.zeros_like(self.ep_return), self.ep_return)
        return obs_out, reward_out, final_out, finished
