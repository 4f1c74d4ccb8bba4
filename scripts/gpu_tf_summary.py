"""Teacher-forced contact-rich summary over all golden tasks (debug)."""
import sys, os, glob
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu import Rig, GOLD, _tasks_with_goldens
for task in (sys.argv[1:] or _tasks_with_goldens()):
    g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
    if "p_actions" not in g: continue
    rig = Rig(torch, task, g["p_rand_vec"]); rig.reset()
    nq, nv = g["p_qpos"].shape[2], g["p_qvel"].shape[2]
    T = g["p_actions"].shape[1]
    wo = wr = wq = wv = 0; bad = 0
    for t in range(0, T - 1):
        st = rig.eng.get_state()
        for k in range(rig.n):
            st[k]["qpos"][:nq] = g["p_qpos"][k, t]; st[k]["qvel"][:nv] = g["p_qvel"][k, t]
            st[k]["mocap_pos"] = g["p_mocap"][k, t]; st[k]["prev_obs"] = g["p_obs"][k, t][:18]
            st[k]["warm"][:] = 0; st[k]["path_len"] = t + 1
        rig.eng.set_state(st)
        o, r, info, _, _ = rig.step(g["p_actions"][:, t + 1])
        s2 = rig.eng.get_state()
        eo = np.abs(o - g["p_obs"][:, t + 1]).max(); er = np.abs(r - g["p_reward"][:, t + 1]).max()
        eq = np.abs(s2["qpos"][:, :nq] - g["p_qpos"][:, t + 1]).max(); ev = np.abs(s2["qvel"][:, :nv] - g["p_qvel"][:, t + 1]).max()
        wo, wr, wq, wv = max(wo, eo), max(wr, er), max(wq, eq), max(wv, ev); bad += eo > 1e-4
    c = rig.eng.counters()
    print(f"{task:28s} obs {wo:.1e} rew {wr:.1e} qpos {wq:.1e} qvel {wv:.1e} bad_steps {bad}/{T-1} dropped {c['contacts_dropped']} iters/pass {c['solver_iters']/c['forward_passes']:.2f}", flush=True)
