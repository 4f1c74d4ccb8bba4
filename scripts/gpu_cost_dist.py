"""Distribution of per-env step cost in the MT50 workload (run under gpurun)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from metaworld_b200.vector_env import MetaWorldVecEnv
from metaworld_b200 import benchmarks as B
names = B.MT50; N = 4096
tasks_all = B.make_tasks(names, False, seed=42)
tasks = [[t for t in tasks_all if t.env_name == n] for n in names]
env = MetaWorldVecEnv(names, tasks, num_envs=N, seed=42, use_one_hot=True, num_tasks=50, max_episode_steps=500)
env.reset(); env.enable_device_sampler()
g = torch.Generator(device=env.device); g.manual_seed(1)
for step in range(1, 401):
    a = torch.rand(N, 4, device=env.device, generator=g) * 2 - 1
    env.engine.set_profiling(step in (5, 50, 100, 200, 300, 400))      # phase timers only on the sampled steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_torch(a); e1.record()
    if step in (5, 50, 100, 200, 300, 400):
        torch.cuda.synchronize()
        P = env.engine.env_profile().astype(np.float64); c = (P[:, 8] - P[:, 12]) / 1e6      # own work (barrier waits excluded)
        em = env.engine.env_model
        q = np.quantile(c, [0.5, 0.9, 0.99, 0.999, 1.0])
        print(f"step {step}: kernel {e0.elapsed_time(e1):.2f} ms | env cost Mcycles mean {c.mean():.2f} p50 {q[0]:.2f} p90 {q[1]:.2f} p99 {q[2]:.2f} p99.9 {q[3]:.2f} max {q[4]:.2f} | balanced {c.sum()/(148*7)/1965*1e3:.2f} ms | slowest CTA {P[:, 8].max()/1965e3:.2f} ms")
        # CTA view: a CTA lasts as long as its slowest env (col 8 = whole-step cycles incl. barrier waits, col 16 = blockIdx);
        # the hardware hands CTAs to SMs in blockIdx order as SMs free up = greedy list scheduling, simulated here
        blk = P[:, 16].astype(np.int64); nb = int(blk.max()) + 1
        dur = np.zeros(nb); np.maximum.at(dur, blk, P[:, 8] / 1965e3)
        sm_t = np.zeros(148)
        for d in dur:
            i = int(np.argmin(sm_t)); sm_t[i] += d
        lpt = np.zeros(148)
        for d in np.sort(dur)[::-1]:
            i = int(np.argmin(lpt)); lpt[i] += d
        print(f"   LPT on the true durations would give {lpt.max():.2f} ms")
        print(f"   CTAs {nb}: sum(CTA time)/148 {dur.sum()/148:.2f} ms | greedy makespan {sm_t.max():.2f} ms (SM finish spread {sm_t.min():.2f}..{sm_t.max():.2f}) | "
              f"own work/(148*7) {c.sum()/(148*7)/1965*1e3:.2f} ms | CTA time min/median/max {dur.min():.2f}/{np.median(dur):.2f}/{dur.max():.2f} ms")
        top = np.argsort(-c)[:8]
        print("   heaviest:", [(names[em[i]], round(c[i], 1)) for i in top])
        bym = sorted(((c[em == m].mean(), names[m]) for m in range(50)), reverse=True)[:6]
        print("   heaviest tasks (mean):", [(n, round(v, 2)) for v, n in bym])
