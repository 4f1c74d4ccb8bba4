"""Where do the heaviest environments of the MT50 workload spend their step?  (run under gpurun)
Per-env phase cycles (mw_set_profiling) at several episode phases, the heaviest CTAs, and the distribution of
contact / constraint-row counts (sizes the small-capacity tier)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from metaworld_b200.vector_env import MetaWorldVecEnv
from metaworld_b200 import benchmarks as B
from metaworld_b200.engine import Engine
names = B.MT50; N = 4096
tasks_all = B.make_tasks(names, False, seed=42)
tasks = [[t for t in tasks_all if t.env_name == n] for n in names]
env = MetaWorldVecEnv(names, tasks, num_envs=N, seed=42, use_one_hot=True, num_tasks=50, max_episode_steps=500)
env.reset(); env.enable_device_sampler()
g = torch.Generator(device=env.device); g.manual_seed(1)
K = Engine.PROFILE_KEYS
marks = (5, 100, 300, 450)
ncon_hist = np.zeros(128, dtype=np.int64); nefc_hist = np.zeros(512, dtype=np.int64)
for step in range(1, 452):
    a = torch.rand(N, 4, device=env.device, generator=g) * 2 - 1
    prof_step = step in marks or step % 10 == 0
    env.engine.set_profiling(prof_step)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_torch(a); e1.record()
    if prof_step:
        torch.cuda.synchronize()
        P = env.engine.env_profile().astype(np.float64)
        np.add.at(ncon_hist, np.minimum(P[:, 14].astype(int), 127), 1); np.add.at(nefc_hist, np.minimum(P[:, 15].astype(int), 511), 1)
    if step in marks:
        em = env.engine.env_model
        c = P[:, 8] / 1e6
        print(f"== step {step}: kernel {e0.elapsed_time(e1):.2f} ms (profiling on) | env Mcyc mean {c.mean():.2f} max {c.max():.2f} | sum/(148*7)/1965MHz = {c.sum()/(148*7)/1965*1e3:.2f} ms")
        tot = P[:, :9].sum(0)
        tot = P[:, :13].sum(0)
        print("   all envs phase share:", {K[i]: round(tot[i] / tot[8], 3) for i in (0, 1, 2, 3, 4, 5, 6, 7, 12)})
        own = (P[:, 8] - P[:, 12]) / 1e6
        print(f"   own work Mcyc mean {own.mean():.2f} p99 {np.quantile(own, 0.99):.2f} max {own.max():.2f} | sum(own)/(148*7)/1965MHz = {own.sum()/(148*7)/1965*1e3:.2f} ms")
        top = np.argsort(-c)[:21:3]
        for i in top:
            row = P[i]
            print(f"   env {i:5d} {names[em[i]]:28s} {c[i]:6.2f} Mcyc (own {(row[8]-row[12])/1e6:5.2f}) cta {int(row[16])} | " + " ".join(f"{K[k]}={row[k]/row[8]:.2f}" for k in (0, 1, 2, 3, 4, 5, 6, 12)) +
                  f" | newton_its={int(row[13])} ncon_max={int(row[14])} nefc_max={int(row[15])} convex={int(row[9])} epa_exp={int(row[10])} gjk_it={int(row[11])}")
        # per-CTA: is the CTA time the max of its warps' own work?  sum over phases of (max over warps) vs max over warps of sum
        bym = sorted(((c[em == m].mean(), c[em == m].max(), names[m]) for m in range(50)), reverse=True)[:8]
        print("   heaviest tasks (mean, max):", [(n, round(v, 2), round(x, 2)) for v, x, n in bym])
cs = np.cumsum(ncon_hist) / ncon_hist.sum(); es = np.cumsum(nefc_hist) / nefc_hist.sum()
print("ncon_max per env-step quantiles:", {q: int(np.searchsorted(cs, q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0)})
print("nefc_max per env-step quantiles:", {q: int(np.searchsorted(es, q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0)})
print("share of env-steps with ncon<=16 and nefc<=88 (proxy, marginal):", float(cs[16]), float(es[88]))
