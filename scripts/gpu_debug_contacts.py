"""Compare device vs oracle contact lists + qacc at one golden policy state (debug)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu import Rig, GOLD
from oracle.tasks import TASKS as OT
from oracle import mjphys as P
np.set_printoptions(precision=6, suppress=True, linewidth=250)
task, k, t = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
rig = Rig(torch, task, g["p_rand_vec"]); rig.reset()
nq, nv = g["p_qpos"].shape[2], g["p_qvel"].shape[2]
oe = OT[task](); rv = g["p_rand_vec"][k]; oe.set_task_vec(rv[:len(oe.random_reset_space()[0])], False); oe.reset()
for sub in range(int(sys.argv[4]) if len(sys.argv) > 4 else 1):
    st = rig.eng.get_state()
    if sub == 0:
        st[k]["qpos"][:nq] = g["p_qpos"][k, t]; st[k]["qvel"][:nv] = g["p_qvel"][k, t]; st[k]["mocap_pos"] = g["p_mocap"][k, t]; st[k]["warm"][:] = 0
        rig.eng.set_state(st)
        oe.data.qpos = g["p_qpos"][k, t]; oe.data.qvel = g["p_qvel"][k, t]; oe.data.mocap_pos[0][:] = g["p_mocap"][k, t]; oe.data.qacc_warmstart = 0
        a = g["p_actions"][k, t + 1]
        # same mocap update as the step
        mp = np.clip(g["p_mocap"][k, t] + np.clip(a[:3], -1, 1) * 0.01, oe.mocap_low, oe.mocap_high)
        st = rig.eng.get_state(); st[k]["mocap_pos"] = mp; rig.eng.set_state(st)
        oe.data.mocap_pos[0][:] = mp; oe.data.mocap_quat[0][:] = [1, 0, 1, 0]
        ctrl = (float(a[3]), -float(a[3])); oe.data.ctrl = ctrl
    con, qacc, meta = rig.eng.debug_forward(ctrl)
    P.mj_forward(oe.model, oe.data)
    print(f'--- substep {sub}: dev ncon {int(meta[k,0])} nefc {int(meta[k,1])} it {int(meta[k,2])} | ora ncon {oe.data.ncon} nefc {oe.data.nefc} it {oe.data.solver_iter}')
    print(' qacc err', np.abs(qacc[k, :nv] - oe.data.qacc).max(), '\n dev', qacc[k, :nv], '\n ora', oe.data.qacc)
    oc = sorted([(c.geom1, c.geom2, c.dist, *c.pos, *list(c.frame)[:3], oe.data.efc_force[c.efc_address] if c.efc_address >= 0 else 0) for c in oe.data.contact])
    dc = sorted([(int(c[7]), int(c[8]), c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[9]) for c in con[k][: int(meta[k, 0])]])
    for x in dc: print('  dev', np.array(x))
    for x in oc: print('  ora', np.array(x))
    rig.eng.debug_substeps(1, ctrl); P.mj_step(oe.model, oe.data, 1)
