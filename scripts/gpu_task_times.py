"""Per-task step time and phase profile (run under gpurun): 888 envs (= 148 SMs x 6 warps) of one task, random actions."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
from metaworld_b200.vector_env import MetaWorldVecEnv
from metaworld_b200 import benchmarks as B
from metaworld_b200.tasks import TASKS
names = sys.argv[1:] or list(TASKS)
import os
N, K = int(os.environ.get("MW_N", 888)), int(os.environ.get("MW_K", 20))
rows = []
for n in names:
    tasks = B.make_tasks([n], False, seed=1, n_goals=10)
    env = MetaWorldVecEnv([n], [tasks], num_envs=N, seed=3, use_one_hot=False, max_episode_steps=500)
    env.reset(); env.enable_device_sampler(); env.engine.set_profiling(True)
    a = torch.rand(K + 3, N, 4, device=env.device) * 2 - 1
    for i in range(3): env.step_torch(a[i])
    torch.cuda.synchronize(); env.engine.profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): env.step_torch(a[3 + i])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    p = env.engine.profile(); c = env.engine.counters()
    st = max(1, p["step"] - p["barrier_wait"])
    rows.append(dict(task=n, ms=ms, mcycles=(p["step"] - p["barrier_wait"]) / (N * K) / 1e6, wait=p["barrier_wait"] / max(1, p["step"] - p["barrier_wait"]), collide=p["collide"] / st, gjk=p["gjk_epa"] / st, solver=p["solver"] / st,
                     pairs=p["n_convex_pairs"] / (N * K), epa=p["n_epa_expansions"] / (N * K), gjkit=p["n_gjk_iters"] / (N * K),
                     newton=c["solver_iters"] / max(1, c["forward_passes"])))
    print(json.dumps(rows[-1]), flush=True)
    env.close() if hasattr(env, "close") else None
