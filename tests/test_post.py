"""metaworld_b200.post.StepPost against scalar transcriptions of the reference wrappers it vectorises
(metaworld/wrappers.py:35-88 RNNBasedMetaRLWrapper, :233-258 NormalizeRewardsExponential; stacking order and
RecordEpisodeStatistics placement from metaworld/__init__.py:437-446)."""
import numpy as np
import pytest

from metaworld_b200.post import StepPost


class ScalarStack:
    """one sub-env's wrapper stack, written the way the reference applies it to a single env"""

    def __init__(self, recurrent, norm_in_obs, exponential, alpha):
        self.recurrent, self.norm_in_obs, self.exponential, self.alpha = recurrent, norm_in_obs, exponential, alpha
        self.mean, self.var, self.ep_ret = 0.0, 1.0, 0.0

    def reset(self, obs):
        self.ep_ret = 0.0
        return np.concatenate([obs, np.zeros(4), [0.0], [0.0]]) if self.recurrent else obs

    def _upd(self, r):
        self.mean = (1 - self.alpha) * self.mean + self.alpha * r
        self.var = (1 - self.alpha) * self.var + self.alpha * np.square(r - self.mean)

    def step(self, next_obs, action, reward, term, trunc):
        o = next_obs
        if self.recurrent:
            o = np.concatenate([next_obs, action, [float(reward) / 10.0 if self.norm_in_obs else float(reward)], [float(term or trunc)]])
        r = reward
        if self.exponential:
            self._upd(reward)          # step(): explicit update ...
            self._upd(reward)          # ... and a second one inside _apply_normalize_reward
            r = reward / (np.sqrt(self.var) + 1e-8)
        self.ep_ret += r
        return o, r


def test_steppost_matches_scalar_wrappers():
    rng = np.random.default_rng(0)
    N, D = 6, 49
    for recurrent, norm_in_obs, expo in [(True, True, True), (True, False, False), (False, True, True), (False, False, False)]:
        post = StepPost(N, recurrent, norm_in_obs, "exponential" if expo else None, reward_alpha=0.05)
        stacks = [ScalarStack(recurrent, norm_in_obs, expo, 0.05) for _ in range(N)]
        obs0 = rng.normal(size=(N, D))
        out = post.on_reset(obs0)
        ref = np.stack([s.reset(obs0[i]) for i, s in enumerate(stacks)])
        assert np.array_equal(out, ref)
        for t in range(40):
            term_obs, reset_obs = rng.normal(size=(N, D)), rng.normal(size=(N, D))
            act, rew = rng.uniform(-1, 1, size=(N, 4)), rng.normal(size=N) * 3
            term, trunc = rng.random(N) < 0.1, rng.random(N) < 0.1
            done = term | trunc
            obs_in = np.where(done[:, None], reset_obs, term_obs)                       # SAME_STEP autoreset: obs is the reset obs
            o, r, fo, fin = post.on_step(obs_in, act, rew, term, trunc, final_obs=term_obs)
            for i, s in enumerate(stacks):
                so, sr = s.step(term_obs[i], act[i], rew[i], term[i], trunc[i])
                assert np.isclose(r[i], sr, rtol=0, atol=1e-12)
                if done[i]:
                    assert np.allclose(fo[i], so) and np.isclose(fin[i], s.ep_ret)
                    assert np.allclose(o[i], s.reset(reset_obs[i]))
                else:
                    assert np.allclose(o[i], so) and fin[i] == 0.0


@pytest.mark.parametrize("cfg", [(True, True, "exponential", 0.05, False), (False, True, "gymnasium", 0.001, True),
                                 (True, False, "gymnasium", 0.001, True), (False, True, None, 0.001, True)])
def test_torch_variant_equals_numpy_variant(cfg):
    """post.StepPostTorch (what step_torch applies on the device) against post.StepPost, on CPU tensors."""
    import torch
    from metaworld_b200.post import StepPost, StepPostTorch
    n, d = 6, 9
    rng = np.random.default_rng(3)
    a = StepPost(n, *cfg)
    b = StepPostTorch(torch, torch.device("cpu"), n, d, *cfg)
    o0 = rng.normal(size=(n, d)).astype(np.float32)
    # (float32 statistics on the numpy side when the input is float32 - the recurrent wrapper's dtype -, float64 on the torch side)
    assert np.allclose(a.on_reset(o0), b.on_reset(torch.from_numpy(o0)).numpy(), atol=1e-6, rtol=1e-3)
    for t in range(12):
        obs = rng.normal(size=(n, d)).astype(np.float32); fo = rng.normal(size=(n, d)).astype(np.float32)
        act = rng.uniform(-1, 1, size=(n, 4)).astype(np.float32); rew = rng.uniform(0, 10, size=n)
        term = rng.random(n) < 0.1; trunc = (t % 5 == 4) & ~term
        x = a.on_step(obs, act, rew, term, trunc, final_obs=fo)
        y = b.on_step(torch.from_numpy(obs), torch.from_numpy(act), torch.from_numpy(rew.astype(np.float32)), torch.from_numpy(term), torch.from_numpy(trunc), torch.from_numpy(fo))
        r32 = rew.astype(np.float32).astype(np.float64)       # the device path sees the float32 reward
        done = term | trunc
        assert np.allclose(x[0], y[0].numpy(), atol=2e-5, rtol=1e-3)
        assert np.allclose(x[2][done], y[2].numpy()[done], atol=2e-5, rtol=1e-3)          # terminal rows are defined for finished envs only
        assert np.allclose(x[1], y[1].numpy(), rtol=1e-5) and np.allclose(x[3], y[3].numpy(), rtol=1e-5, atol=1e-6)
