"""Vectorised equivalents of the two per-sub-env wrappers the reference stacks between the one-hot wrapper and the
episode-statistics wrapper (metaworld/__init__.py:437-444):

* `RNNBasedMetaRLWrapper` (metaworld/wrappers.py:35-88): obs <- [obs, action, reward (/10), done]; after a reset the three
  extra blocks are zero;
* `NormalizeRewardsExponential` (wrappers.py:233-258): per-env exponential running mean / variance of the reward; the
  reference updates the estimate twice per step (once in `step`, once more inside `_apply_normalize_reward`), which is
  reproduced here because it changes the numbers.

Pure numpy on [N, ...] arrays; `MetaWorldVecEnv` applies it to what the engine returns.  Because the reference puts
`RecordEpisodeStatistics` outside the reward normalisation, the episodic return it reports is the sum of NORMALISED
rewards; `ep_return` below tracks that."""
from __future__ import annotations

import numpy as np


class StepPost:
    def __init__(self, num_envs, recurrent_info_in_obs=False, normalize_reward_in_recurrent_info=True,
                 reward_normalization_method=None, reward_alpha=0.001):
        if reward_normalization_method not in (None, "exponential"):
            raise NotImplementedError("reward_normalization_method='gymnasium' relies on gymnasium.wrappers.NormalizeReward; "
                                      "only None and 'exponential' are provided")
        self.n = num_envs
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
        return self.recurrent or self.exponential

    def on_reset(self, obs, mask=None):
        """obs [N, D] -> [N, D + extra]; the reward statistics are NOT reset (the wrapper object lives across episodes)."""
        if mask is None:
            self.ep_return[:] = 0
        else:
            self.ep_return[mask] = 0
        if not self.recurrent:
            return obs
        return np.concatenate([obs, np.zeros((len(obs), 6), dtype=obs.dtype)], axis=1)

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
        self.ep_return += reward_out
        finished = np.where(done, self.ep_return, 0.0)
        self.ep_return[done] = 0
        return obs_out, reward_out, final_out, finished


class StepPostTorch:
    """The same two wrappers on CUDA tensors, for `MetaWorldVecEnv.step_torch` (so RL code that keeps everything on the GPU
    gets the recurrent observation and the normalised reward without a host round trip).  Plain torch elementwise ops on
    [N, ...] tensors: plumbing around the engine's outputs, not a hot path."""

    def __init__(self, torch, device, num_envs, obs_dim, recurrent_info_in_obs=False, normalize_reward_in_recurrent_info=True,
                 reward_normalization_method=None, reward_alpha=0.001):
        self.t = torch
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
        self.mean.copy_(self.t.from_numpy(post.mean)); self.var.copy_(self.t.from_numpy(post.var)); self.ep_return.copy_(self.t.from_numpy(post.ep_return))

    def on_reset(self, obs):
        self.ep_return.zero_()
        if not self.recurrent:
            return obs
        self.out.zero_(); self.out[:, : self.obs_dim] = obs
        return self.out

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
        self.ep_return += reward_out
        finished = t.where(done, self.ep_return, t.zeros_like(self.ep_return))
        self.ep_return = t.where(done, t.zeros_like(self.ep_return), self.ep_return)
        return obs_out, reward_out, final_out, finished
