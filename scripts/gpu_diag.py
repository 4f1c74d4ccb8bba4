"""Debug tool (run under gpurun): device vs oracle, per task.
  python scripts/gpu_diag.py reset  task...   # post-reset obs / qpos / qvel differences per goal
  python scripts/gpu_diag.py open   task...   # first diverging step of the open-loop golden rollout + contact comparison there
  python scripts/gpu_diag.py cr     task...   # worst teacher-forced step on the policy trajectory + contact comparison there
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu import Rig, GOLD  # noqa: E402
from oracle.tasks import TASKS as OT  # noqa: E402
from oracle import mjphys as P  # noqa: E402

np.set_printoptions(precision=6, suppress=True, linewidth=250)


def big(a, b, tol=1e-5):
    d = np.abs(np.asarray(a, float) - np.asarray(b, float))
    idx = np.where(~(d < tol))[0]
    return ", ".join(f"[{i}] dev {a[i]:.6f} ora {b[i]:.6f}" for i in idx[:12])


def contacts(rig, task, rv, k, qpos, qvel, mocap, action, nsub=5):
    """Set both sides to the same state, apply the action's mocap/ctrl, compare forward passes substep by substep."""
    oe = OT[task](); n = len(oe.random_reset_space()[0]); oe.set_task_vec(rv[:n], False); oe.reset()
    nq, nv = len(qpos), len(qvel)
    st = rig.eng.get_state()
    mp = np.clip(mocap + np.clip(action[:3], -1, 1) * 0.01, oe.mocap_low, oe.mocap_high)
    st[k]["qpos"][:nq] = qpos; st[k]["qvel"][:nv] = qvel; st[k]["mocap_pos"] = mp; st[k]["warm"][:] = 0
    rig.eng.set_state(st)
    oe.data.qpos = qpos; oe.data.qvel = qvel; oe.data.mocap_pos[0][:] = mp; oe.data.mocap_quat[0][:] = [1, 0, 1, 0]; oe.data.qacc_warmstart = 0
    ctrl = (float(action[3]), -float(action[3])); oe.data.ctrl = ctrl
    for sub in range(nsub):
        con, qacc, meta = rig.eng.debug_forward(ctrl)
        P.mj_forward(oe.model, oe.data)
        qe = np.abs(qacc[k, :nv] - oe.data.qacc).max()
        print(f'  substep {sub}: dev ncon {int(meta[k,0])} nefc {int(meta[k,1])} it {int(meta[k,2])} | ora ncon {oe.data.ncon} nefc {oe.data.nefc} it {oe.data.solver_iter} | qacc err {qe:.3e}')
        if qe > 1e-2 * max(1.0, np.abs(oe.data.qacc).max()) or not np.isfinite(qe):
            oc = sorted([(c.geom1, c.geom2, c.dist, *c.pos, *list(c.frame)[:3], oe.data.efc_force[c.efc_address] if c.efc_address >= 0 else 0) for c in oe.data.contact])
            dc = sorted([(int(c[7]), int(c[8]), c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[9]) for c in con[k][: int(meta[k, 0])]])
            for x in dc: print('    dev', np.array(x))
            for x in oc: print('    ora', np.array(x))
            print('    dev qacc', qacc[k, :nv]); print('    ora qacc', oe.data.qacc)
            st = rig.eng.get_state(); print('    dev qpos', st[k]["qpos"][:nq]); print('    ora qpos', oe.data.qpos)
            break
        rig.eng.debug_substeps(1, ctrl); P.mj_step(oe.model, oe.data, 1)


def do_first(task):
    """first env step after reset, substep by substep (device snapshot state vs oracle reset state)"""
    g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
    rig = Rig(torch, task, g["rand_vec"]); rig.reset()
    k = 0
    st = rig.eng.get_state()
    nq, nv = g["reset_qpos"].shape[1], g["reset_qvel"].shape[1]
    oe = OT[task](); n = len(oe.random_reset_space()[0]); oe.set_task_vec(g["rand_vec"][k][:n], False); oe.reset()
    print(f"{task}: reset state: qpos {big(st[k]['qpos'][:nq], oe.data.qpos, 1e-7)} | qvel {big(st[k]['qvel'][:nv], oe.data.qvel, 1e-6)} | warm {big(st[k]['warm'][:nv], oe.data.qacc_warmstart, 1e-4)} | mocap {big(st[k]['mocap_pos'], oe.data.mocap_pos[0], 1e-7)}")
    a = g["actions"][k, 0]
    mp = np.clip(oe.data.mocap_pos[0] + np.clip(a[:3], -1, 1) * 0.01, oe.mocap_low, oe.mocap_high)
    st[k]["mocap_pos"] = mp; rig.eng.set_state(st)
    oe.data.mocap_pos[0][:] = mp; oe.data.mocap_quat[0][:] = [1, 0, 1, 0]
    ctrl = (float(a[3]), -float(a[3])); oe.data.ctrl = ctrl
    for sub in range(5):
        con, qacc, meta = rig.eng.debug_forward(ctrl)
        P.mj_forward(oe.model, oe.data)
        qe = np.abs(qacc[k, :nv] - oe.data.qacc).max()
        print(f'  substep {sub}: dev ncon {int(meta[k,0])} nefc {int(meta[k,1])} it {int(meta[k,2])} | ora ncon {oe.data.ncon} nefc {oe.data.nefc} it {oe.data.solver_iter} | qacc err {qe:.3e}')
        oc = sorted([(c.geom1, c.geom2, c.dist, *c.pos, *list(c.frame)[:3], oe.data.efc_force[c.efc_address] if c.efc_address >= 0 else 0) for c in oe.data.contact])
        dc = sorted([(int(c[7]), int(c[8]), c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[9]) for c in con[k][: int(meta[k, 0])]])
        for x in dc: print('    dev', np.array(x))
        for x in oc: print('    ora', np.array(x))
        print('    dev qacc', qacc[k, :nv]); print('    ora qacc', oe.data.qacc)
        rig.eng.debug_substeps(1, ctrl); P.mj_step(oe.model, oe.data, 1)
        st2 = rig.eng.get_state(); print('    qpos err', big(st2[k]['qpos'][:nq], oe.data.qpos, 1e-6))


def do_reset(task):
    g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
    rig = Rig(torch, task, g["rand_vec"]); snaps = rig.eng.get_snapshots()
    nq, nv = g["reset_qpos"].shape[1], g["reset_qvel"].shape[1]
    for k in range(rig.n):
        print(f"{task} goal {k}: obs: {big(snaps[k]['obs'], g['reset_obs'][k])} | qpos: {big(snaps[k]['st']['qpos'][:nq], g['reset_qpos'][k])} | qvel: {big(snaps[k]['st']['qvel'][:nv], g['reset_qvel'][k], 1e-4)}")


def do_open(task):
    g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
    rv = g["rand_vec"].astype(np.float32) if os.environ.get("MW_DIAG_RV32") else g["rand_vec"]     # (float32-rounded goals: the pre-0892805 behaviour)
    rig = Rig(torch, task, rv); rig.reset()
    for t in range(g["actions"].shape[1]):
        o, r, info, _, _ = rig.step(g["actions"][:, t])
        e = np.abs(o - g["obs"][:, t]).max(axis=1)
        print(f"   t {t} obs err per env {e} reward err {np.abs(r - g['reward'][:, t])}")
        if (e > 1e-4).any() or not np.isfinite(e).all():
            k = int(np.nanargmax(np.where(np.isfinite(e), e, 1e9)))
            print(f"{task}: open-loop first divergence at step {t} env {k}: {big(o[k], g['obs'][k, t], 1e-4)}")
            if t > 0:
                contacts(rig, task, g["rand_vec"][k], k, g["qpos"][k, t - 1], g["qvel"][k, t - 1], g["mocap"][k, t - 1], g["actions"][k, t])
            return
    print(f"{task}: open-loop clean")


def do_cr(task):
    g = np.load(os.path.join(GOLD, f"traj_{task}.npz"))
    rig = Rig(torch, task, g["p_rand_vec"]); rig.reset()
    rig.step(g["p_actions"][:, 0])     # like the test: rewards that latch state on their first call latch here
    nq, nv = g["p_qpos"].shape[2], g["p_qvel"].shape[2]
    worst = (-1, 0, 0)
    for t in range(g["p_actions"].shape[1] - 1):
        st = rig.eng.get_state()
        for k in range(rig.n):
            st[k]["qpos"][:nq] = g["p_qpos"][k, t]; st[k]["qvel"][:nv] = g["p_qvel"][k, t]
            st[k]["mocap_pos"] = g["p_mocap"][k, t]; st[k]["prev_obs"] = g["p_obs"][k, t][:18]; st[k]["warm"][:] = 0; st[k]["path_len"] = t + 1
        rig.eng.set_state(st)
        o, r, info, _, _ = rig.step(g["p_actions"][:, t + 1])
        e = np.abs(o - g["p_obs"][:, t + 1]).max(axis=1)
        e = np.where(np.isfinite(e), e, 1e9)
        if e.max() > worst[0]:
            worst = (float(e.max()), t, int(e.argmax()))
    e, t, k = worst
    print(f"{task}: contact-rich worst step err {e:.3e} at t {t} env {k}")
    contacts(rig, task, g["p_rand_vec"][k], k, g["p_qpos"][k, t], g["p_qvel"][k, t], g["p_mocap"][k, t], g["p_actions"][k, t + 1])


if __name__ == "__main__":
    mode = sys.argv[1]
    for task in sys.argv[2:]:
        try:
            dict(reset=do_reset, open=do_open, cr=do_cr, first=do_first)[mode](task)
        except Exception as ex:  # keep going: this is a survey tool
            print(f"{task}: {mode} raised {ex!r}")
