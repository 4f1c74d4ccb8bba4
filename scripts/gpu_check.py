"""Ad-hoc GPU check: device engine vs CPU oracle for one task (run under gpurun)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from metaworld_b200.engine import Engine, lib
from metaworld_b200.tasks import TASKS
from oracle.tasks import TASKS as OT

task = sys.argv[1] if len(sys.argv) > 1 else 'reach-v3'
np.set_printoptions(precision=6, suppress=True, linewidth=220)
print(lib().mw_build_info().decode())
spec = TASKS[task]
eng = Engine([task])
rng = np.random.default_rng(0)
K = 4
rvs = rng.uniform(spec.rand_low, spec.rand_high, size=(K, 6))
t = time.time()
ids = eng.build_snapshots([0] * K, rvs, [0] * K)
print('snapshot build', time.time() - t, 's')
snaps = eng.get_snapshots()
# oracle resets
maxerr = 0
oenvs = []
for k in range(K):
    oe = OT[task](); oe.set_task_vec(rvs[k], False)
    oobs, _ = oe.reset()
    oenvs.append(oe)
    nq, nv = oe.model.nq, oe.model.nv
    dq = np.abs(snaps[k]['st']['qpos'][:nq] - oe.data.qpos).max()
    dv = np.abs(snaps[k]['st']['qvel'][:nv] - oe.data.qvel).max()
    do = np.abs(snaps[k]['obs'] - oobs).max()
    print(f'snap {k}: dq {dq:.2e} dv {dv:.2e} dobs {do:.2e}')
    if k == 0:
        print(' dev qpos', snaps[k]['st']['qpos'][:nq]); print(' ora qpos', oe.data.qpos)
        print(' dev obs', snaps[k]['obs']); print(' ora obs', oobs)
# step parity, open loop
N = K
eng.set_envs([0] * N)
eng.set_options(500, False, 0)
dev = eng.device
obs = torch.zeros(N, 39, device=dev); rew = torch.zeros(N, device=dev)
term = torch.zeros(N, dtype=torch.uint8, device=dev); trunc = torch.zeros(N, dtype=torch.uint8, device=dev)
info = torch.zeros(N, 7, device=dev); fobs = torch.zeros(N, 39, device=dev); finfo = torch.zeros(N, 8, device=dev)
sid = torch.tensor(ids, dtype=torch.int32, device=dev)
eng.reset(sid, obs)
torch.cuda.synchronize()
print('reset obs err', np.abs(obs.cpu().numpy() - np.stack([s['obs'] for s in snaps])).max())
arng = np.random.default_rng(1)
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for t_ in range(T):
    a = arng.uniform(-1, 1, size=(N, 4)).astype(np.float32)
    eng.step(torch.tensor(a, device=dev), obs, rew, term, trunc, info, fobs, finfo, sid)
    torch.cuda.synchronize()
    o = obs.cpu().numpy(); r = rew.cpu().numpy(); inf = info.cpu().numpy()
    eo = []; er = []
    for k in range(N):
        oo, rr, _, tr, ii = oenvs[k].step(a[k])
        eo.append(np.abs(oo - o[k]).max()); er.append(abs(rr - r[k]))
    if t_ % 10 == 0 or t_ == T - 1:
        print(t_, 'obs err', np.max(eo), 'rew err', np.max(er), 'ncon/iters', eng.counters())
st = eng.get_state()
print('final dq', np.abs(st[0]['qpos'][:nq] - oenvs[0].data.qpos).max())
# timing
N = 4096
eng.set_envs([0] * N)
sid = torch.tensor(np.resize(ids, N), dtype=torch.int32, device=dev)
obs = torch.zeros(N, 39, device=dev); rew = torch.zeros(N, device=dev)
term = torch.zeros(N, dtype=torch.uint8, device=dev); trunc = torch.zeros(N, dtype=torch.uint8, device=dev)
info = torch.zeros(N, 7, device=dev); fobs = torch.zeros(N, 39, device=dev); finfo = torch.zeros(N, 8, device=dev)
eng.reset(sid, obs)
act = torch.rand(N, 4, device=dev) * 2 - 1
for _ in range(3): eng.step(act, obs, rew, term, trunc, info, fobs, finfo, sid)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
S = 50
for _ in range(S):
    act = torch.rand(N, 4, device=dev) * 2 - 1
    eng.step(act, obs, rew, term, trunc, info, fobs, finfo, sid)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'{N} envs: {ms / S:.3f} ms/step -> {N * S / ms * 1e3:.0f} env steps/s', eng.counters())
