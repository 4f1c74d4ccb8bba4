"""Where the numpy-API step spends host time beyond the kernel (MT50 @ 4096, steady state).  Run under gpurun."""
import sys, time, cProfile, pstats, io, numpy as np, torch
sys.path.insert(0, '.')
from bench import build_env, stagger
from types import SimpleNamespace
env, names, _, _ = build_env(SimpleNamespace(benchmark="MT50", envs_per_gpu=4096, seed=42), 0, 0)
N = env.num_envs
env.reset(); stagger(env, 0)
rng = np.random.default_rng(1)
acts = [rng.uniform(-1, 1, size=(N, 4)).astype(np.float32) for _ in range(16)]
for i in range(30): env.step(acts[i % 16])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(100): env.step(acts[i % 16])
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"numpy step: {(t1 - t0) * 10:.3f} ms / step")
pr = cProfile.Profile(); pr.enable()
for i in range(100): env.step(acts[i % 16])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25); print(s.getvalue())
# kernel alone
env.enable_device_sampler()
a = torch.from_numpy(acts[0]).to(env.device)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(10): env.step_torch(a)
e0.record()
for i in range(100): env.step_torch(a)
e1.record(); torch.cuda.synchronize()
print(f"step_torch (device timed): {e0.elapsed_time(e1) / 100:.3f} ms / step")
