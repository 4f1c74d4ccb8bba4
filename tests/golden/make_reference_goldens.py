"""Regenerates tests/golden/traj_<task>.npz from the REFERENCE's own env classes.

"Reference glue on restated physics": `/root/reference/metaworld/sawyer_xyz_env.py` and
`metaworld/envs/sawyer_*_v3.py` run UNMODIFIED on `oracle/refshim` (stand-ins for the uninstallable `mujoco` /
`gymnasium`, see oracle/refshim/README.md), whose `mj_step` / `mj_forward` are this repo's float64 restatement
(oracle/mjphys.c).  Every line of reference Python above the `mj_*` calls is therefore the reference's; the physics
below them stays PARITY UNPINNED against real MuJoCo.

Per task the file holds
  * 3 goals x 60 random-action steps (seed 42 goals, action rng 2024): reset obs/qpos/qvel, per-step obs, reward,
    all 7 info keys, qpos, qvel, mocap;
  * 2 goals x 120 steps driven by the reference's scripted policy (contact-rich), keys prefixed `p_`;
  * 1 goal x 60 steps with `partially_observable=True` (ML benchmarks), keys prefixed `po_`;
  * 1 goal x 500 random-action steps (a full episode; the last step has truncate=True), keys prefixed `l_`;
  * 5 goals x the reference's scripted policy run closed-loop until success (+5 steps, at most 300): only the action
    sequence, its length and whether the reference run succeeded, keys prefixed `s_` (replayed open-loop on the device,
    where the reference's policy code is not available: tests/test_gpu.py::test_scripted_policy_actions_succeed).

Run here (needs /root/reference):  python tests/golden/make_reference_goldens.py [task ...]
tests/test_refpin.py checks oracle/tasks.py + oracle/sawyer_env.py against these files to 1e-12.
"""
import os
import pickle
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from oracle import refshim  # noqa: E402

KEYS = ("success", "near_object", "grasp_success", "grasp_reward", "in_place_reward", "obj_to_target", "unscaled_reward")


def reference_env(task, rand_vec, partially_observable=False):
    refshim.activate()
    from metaworld.env_dict import ALL_V3_ENVIRONMENTS
    from metaworld.types import Task

    cls = ALL_V3_ENVIRONMENTS[task]
    env = cls()
    env.set_task(Task(env_name=task, data=pickle.dumps(
        dict(rand_vec=np.asarray(rand_vec, dtype=np.float64), env_cls=cls, partially_observable=partially_observable))))
    return env


def reference_policy(task):
    refshim.activate()
    from metaworld.policies import ENV_POLICY_MAP

    return ENV_POLICY_MAP[task]()


def rollout(env, actions=None, policy=None, T=None):
    o, _ = env.reset()
    first = dict(reset_obs=o.copy(), reset_qpos=env.data.qpos.copy(), reset_qvel=env.data.qvel.copy())
    tr = dict(actions=[], obs=[], reward=[], info=[], qpos=[], qvel=[], mocap=[], truncated=[])
    for t in range(T if actions is None else len(actions)):
        a = (np.clip(policy.get_action(o.copy()), -1, 1).astype(np.float32) if actions is None else actions[t])
        o, r, term, trunc, info = env.step(a)
        tr["actions"].append(a); tr["obs"].append(o); tr["reward"].append(float(r)); tr["truncated"].append(bool(trunc))
        tr["info"].append([float(info[k]) for k in KEYS])
        tr["qpos"].append(env.data.qpos.copy()); tr["qvel"].append(env.data.qvel.copy()); tr["mocap"].append(env.data.mocap_pos[0].copy())
    return first, {k: np.array(v) for k, v in tr.items()}


def random_actions(rng, T):
    A = rng.uniform(-1, 1, size=(T, 4)).astype(np.float32)
    A[T // 2:, 3] = 1.0        # second half closes the gripper (exercises pad contacts / grasp branches)
    return A


def make(task, factory=reference_env, policy_factory=reference_policy):
    from metaworld_b200 import benchmarks as B

    def pad(rv):
        return np.pad(np.asarray(rv, dtype=np.float64), (0, 6 - len(rv)))

    out = {}

    def add(prefix, rows):
        for row in rows:
            for k, v in row.items():
                out.setdefault(prefix + k, []).append(v)

    rng = np.random.default_rng(2024)
    rows = []
    for tk in B.make_tasks([task], False, seed=42, n_goals=3):
        rv = tk.unpack()["rand_vec"]
        first, tr = rollout(factory(task, rv), actions=random_actions(rng, 60))
        rows.append(dict(rand_vec=pad(rv), **first, **tr))
    add("", rows)
    out["success"] = [r["info"][:, 0] for r in rows]

    rows = []
    for tk in B.make_tasks([task], False, seed=7, n_goals=2):
        rv = tk.unpack()["rand_vec"]
        first, tr = rollout(factory(task, rv), policy=policy_factory(task), T=120)
        rows.append(dict(rand_vec=pad(rv), len=120, **first, **tr))
    add("p_", rows)

    rv = B.make_tasks([task], False, seed=11, n_goals=1)[0].unpack()["rand_vec"]
    first, tr = rollout(factory(task, rv, True), actions=random_actions(np.random.default_rng(77), 60))
    add("po_", [dict(rand_vec=pad(rv), **first, **tr)])

    rv = B.make_tasks([task], False, seed=13, n_goals=1)[0].unpack()["rand_vec"]
    A = np.random.default_rng(500).uniform(-1, 1, size=(500, 4)).astype(np.float32)
    first, tr = rollout(factory(task, rv), actions=A)
    add("l_", [dict(rand_vec=pad(rv), **first, **tr)])

    acts, lens, succ, rvs = [], [], [], []
    for tk in B.make_tasks([task], False, seed=21, n_goals=5):
        rv = tk.unpack()["rand_vec"]
        env, pol = factory(task, rv), policy_factory(task)
        o, _ = env.reset()
        A = np.zeros((300, 4), dtype=np.float32)
        n, ok, stop = 0, False, 300
        while n < stop:
            a = np.clip(pol.get_action(o.copy()), -1, 1).astype(np.float32)
            o, r, _, _, info = env.step(a)
            A[n] = a; n += 1
            if info["success"] == 1.0 and not ok:
                ok, stop = True, min(300, n + 5)
        acts.append(A); lens.append(n); succ.append(ok); rvs.append(pad(rv))
    out["s_actions"], out["s_len"], out["s_success"], out["s_rand_vec"] = acts, lens, succ, rvs

    out["source"] = np.frombuffer(b"reference glue (metaworld 3.1.1 classes, unmodified) on restated physics (oracle/mjphys)", dtype=np.uint8)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"traj_{task}.npz"),
                        **{k: np.array(v) for k, v in out.items()})
    print(task, "ok", flush=True)


if __name__ == "__main__":
    from metaworld_b200.tasks import TASKS as _T

    names = sys.argv[1:] or list(_T)
    if len(names) > 4:
        import multiprocessing as mp

        with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
            pool.map(make, names)
    else:
        for t in names:
            make(t)
