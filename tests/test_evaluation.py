"""`metaworld_b200.evaluation` against a literal transcription of the reference's per-env bookkeeping loop
(metaworld/evaluation.py:60-104), driven by a scripted fake vector env (no GPU)."""
import numpy as np

from metaworld_b200 import evaluation as E


class FakeVec:
    """Vector-env protocol stub: env i of task names[i]; episodes end on a scripted schedule with scripted returns / success."""

    def __init__(self, names, seed=0):
        self.names, self.num_envs = names, len(names)
        self.rng = np.random.default_rng(seed)
        self.tos = False
        self.calls = []

    def get_attr(self, name):
        if name == "terminate_on_success":
            return tuple([self.tos] * self.num_envs)
        if name == "task_name":
            return tuple(self.names)
        raise KeyError(name)

    def call(self, name, *a):
        self.calls.append((name, a))
        if name == "toggle_terminate_on_success":
            self.tos = bool(a[0])

    def reset(self):
        return np.zeros((self.num_envs, 39)), {}

    def step(self, actions):
        n = self.num_envs
        term = self.rng.random(n) < 0.2
        trunc = (~term) & (self.rng.random(n) < 0.1)
        done = term | trunc
        info = {}
        if done.any():
            info["final_info"] = {"episode": {"r": np.where(done, self.rng.normal(size=n) * 10, 0.0)},
                                  "success": np.where(done, (self.rng.random(n) < 0.5).astype(float), 0.0)}
        return self.rng.normal(size=(n, 39)), self.rng.normal(size=n), term, trunc, info


class NullAgent:
    def __init__(self): self.resets = 0
    def eval_action(self, obs): return np.zeros((len(obs), 4))
    def reset(self, mask): self.resets += int(np.sum(mask))


def reference_loop(agent, envs, num_episodes):
    """transcription of the reference's loop, kept deliberately per-env"""
    envs.call("toggle_terminate_on_success", True)
    obs, _ = envs.reset()
    agent.reset(np.ones(envs.num_envs, dtype=bool))
    task_names = list(envs.get_attr("task_name"))
    successes = {t: 0 for t in set(task_names)}
    rets = {t: [] for t in set(task_names)}
    while not all(len(r) >= num_episodes for r in rets.values()):
        obs, _, term, trunc, infos = envs.step(agent.eval_action(obs))
        dones = np.logical_or(term, trunc)
        agent.reset(dones)
        for i, ended in enumerate(dones):
            if ended:
                rets[task_names[i]].append(float(infos["final_info"]["episode"]["r"][i]))
                if len(rets[task_names[i]]) <= num_episodes:
                    successes[task_names[i]] += int(infos["final_info"]["success"][i])
    rets = {t: r[:num_episodes] for t, r in rets.items()}
    sr = {t: s / num_episodes for t, s in successes.items()}
    return float(np.mean(list(sr.values()))), float(np.mean(list(rets.values()))), sr, rets


def test_evaluation_matches_reference_bookkeeping():
    names = ["reach-v3"] * 5 + ["push-v3"] * 3 + ["door-open-v3"] * 4
    for seed in range(5):
        a = E.evaluation(NullAgent(), FakeVec(names, seed), num_episodes=7)
        b = reference_loop(NullAgent(), FakeVec(names, seed), 7)
        assert a[0] == b[0] and abs(a[1] - b[1]) < 1e-12 and a[2] == b[2] and a[3] == b[3]   # mean over tasks: summation order only


def test_evaluation_restores_terminate_on_success():
    v = FakeVec(["reach-v3"] * 4)
    E.evaluation(NullAgent(), v, num_episodes=2)
    assert v.tos is False and v.calls[0] == ("toggle_terminate_on_success", (True,))


class MetaAgent(NullAgent):
    def __init__(self): super().__init__(); self.inits = self.adapts = self.steps = 0
    def init(self): self.inits += 1
    def adapt_action(self, obs): return np.zeros((len(obs), 4)), {"x": np.zeros(len(obs))}
    def step(self, ts): self.steps += 1; assert isinstance(ts, E.Timestep)
    def adapt(self): self.adapts += 1


def test_metalearning_evaluation_protocol():
    v = FakeVec(["reach-v3"] * 3 + ["push-v3"] * 3, seed=3)
    ag = MetaAgent()
    sr, ret, per_task = E.metalearning_evaluation(ag, v, num_evals=2, adaptation_steps=2, adaptation_episodes=2, evaluation_episodes=2)
    assert ag.inits == 2 and ag.adapts == 4 and ag.steps > 0
    assert set(per_task) == {"reach-v3", "push-v3"} and 0.0 <= sr <= 1.0
    assert [c[0] for c in v.calls[:3]] == ["toggle_sample_tasks_on_reset", "toggle_terminate_on_success", "sample_tasks"]
