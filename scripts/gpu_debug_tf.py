"""Teacher-forced per-step comparison along the policy trajectory of a golden file (debug)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_gpu import Rig, GOLD
np.set_printoptions(precision=6, suppress=True, linewidth=250)
task = sys.argv[1]
g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
rig = Rig(torch, task, g["p_rand_vec"]); rig.reset()
nq, nv = g["p_qpos"].shape[2], g["p_qvel"].shape[2]
T = g["p_actions"].shape[1]
for t in range(0, T - 1):
    st = rig.eng.get_state()
    for k in range(rig.n):
        st[k]["qpos"][:nq] = g["p_qpos"][k, t]; st[k]["qvel"][:nv] = g["p_qvel"][k, t]
        st[k]["mocap_pos"] = g["p_mocap"][k, t]; st[k]["prev_obs"] = g["p_obs"][k, t][:18]
        st[k]["warm"][:] = 0; st[k]["path_len"] = t + 1
    rig.eng.set_state(st)
    o, r, info, _, _ = rig.step(g["p_actions"][:, t + 1])
    s2 = rig.eng.get_state()
    eo = np.abs(o - g["p_obs"][:, t + 1]).max(axis=1); er = np.abs(r - g["p_reward"][:, t + 1])
    eq = np.abs(s2["qpos"][:, :nq] - g["p_qpos"][:, t + 1]).max(axis=1); ev = np.abs(s2["qvel"][:, :nv] - g["p_qvel"][:, t + 1]).max(axis=1)
    flag = '  <<<' if max(eo.max(), er.max()) > 1e-4 else ''
    print(t, 'obs', eo, 'rew', er, 'qpos', eq, 'qvel', ev, flag)
    if flag and '-v' in sys.argv:
        k = int(np.argmax(eo + er))
        print('   dev obs', o[k][:18]); print('   ora obs', g["p_obs"][k, t + 1][:18]); print('   rew', r[k], g["p_reward"][k, t+1], 'info', info[k], g["p_info"][k, t+1])
        print('   dq', s2["qpos"][k, :nq] - g["p_qpos"][k, t + 1]); print('   dv', s2["qvel"][k, :nv] - g["p_qvel"][k, t + 1])
print(rig.eng.counters())
