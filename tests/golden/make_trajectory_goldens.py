"""Generates tests/golden/traj_<task>.npz with the CPU oracle (oracle/): for seeded goals and a seeded random
action sequence, the post-reset state/obs and the per-step (obs, reward, success, qpos, qvel, mocap).
PARITY UNPINNED: the oracle is a restatement (MuJoCo is not installable here), so these are oracle goldens,
not MuJoCo goldens.  Run here:  python tests/golden/make_trajectory_goldens.py [task ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.tasks import TASKS as OT
from metaworld_b200 import benchmarks as B

def _policy(task):
    """Reference scripted policy for `task`, or None when /root/reference is not available."""
    import types
    ref = "/root/reference/metaworld"
    if not os.path.isdir(ref):
        return None
    if "metaworld" not in sys.modules:
        pkg = types.ModuleType("metaworld"); pkg.__path__ = [ref]; sys.modules["metaworld"] = pkg
    import warnings
    warnings.simplefilter("ignore")
    import metaworld.policies as MP
    return MP.ENV_POLICY_MAP[task]()


def make(task, n_goals=3, T=60):
    tasks = B.make_tasks([task], False, seed=42, n_goals=n_goals)
    rng = np.random.default_rng(2024)
    out = dict(rand_vec=[], reset_obs=[], reset_qpos=[], reset_qvel=[], actions=[], obs=[], reward=[], success=[], qpos=[], qvel=[], mocap=[], info=[])
    for tk in tasks:
        env = OT[task]()
        rv = tk.unpack()["rand_vec"]
        env.set_task_vec(rv, False)
        o, _ = env.reset()
        out["rand_vec"].append(np.pad(rv, (0, 6 - len(rv)))); out["reset_obs"].append(o)
        out["reset_qpos"].append(env.data.qpos.copy()); out["reset_qvel"].append(env.data.qvel.copy())
        A = rng.uniform(-1, 1, size=(T, 4)).astype(np.float32)
        # mix in a few saturated / gripper-closing actions
        A[T // 2:, 3] = 1.0
        tr = dict(obs=[], reward=[], success=[], qpos=[], qvel=[], mocap=[], info=[])
        for a in A:
            o, r, _, _, info = env.step(a)
            tr["obs"].append(o); tr["reward"].append(r); tr["success"].append(info["success"])
            tr["qpos"].append(env.data.qpos.copy()); tr["qvel"].append(env.data.qvel.copy()); tr["mocap"].append(env.data.mocap_pos[0].copy())
            tr["info"].append([float(info[k]) for k in ("success", "near_object", "grasp_success", "grasp_reward", "in_place_reward", "obj_to_target", "unscaled_reward")])
        out["actions"].append(A)
        for k in tr: out[k].append(np.array(tr[k]))
    # contact-rich trajectories: the REFERENCE's scripted policy (metaworld/policies, numpy only) drives the oracle
    pol = _policy(task)
    if pol is not None:
        P = dict(p_rand_vec=[], p_actions=[], p_obs=[], p_reward=[], p_info=[], p_qpos=[], p_qvel=[], p_mocap=[], p_len=[])
        for tk in B.make_tasks([task], False, seed=7, n_goals=2):
            env = OT[task]()
            rv = tk.unpack()["rand_vec"]
            env.set_task_vec(rv, False)
            o, _ = env.reset()
            tr = dict(p_actions=[], p_obs=[], p_reward=[], p_info=[], p_qpos=[], p_qvel=[], p_mocap=[])
            TP = 120
            for t in range(TP):
                a = np.clip(pol.get_action(o.copy()), -1, 1).astype(np.float32)   # policies mutate their input
                o, r, _, _, info = env.step(a)
                tr["p_actions"].append(a); tr["p_obs"].append(o); tr["p_reward"].append(r)
                tr["p_qpos"].append(env.data.qpos.copy()); tr["p_qvel"].append(env.data.qvel.copy()); tr["p_mocap"].append(env.data.mocap_pos[0].copy())
                tr["p_info"].append([float(info[k]) for k in ("success", "near_object", "grasp_success", "grasp_reward", "in_place_reward", "obj_to_target", "unscaled_reward")])
            P["p_rand_vec"].append(np.pad(rv, (0, 6 - len(rv)))); P["p_len"].append(TP)
            for k in tr: P[k].append(np.array(tr[k]))
        out.update(P)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), f"traj_{task}.npz"), **{k: np.array(v) for k, v in out.items()})
    print(task, "ok")

if __name__ == "__main__":
    for t in (sys.argv[1:] or list(OT)):
        make(t)
