"""ncu target: MT50 @ 4096 in steady state (episode phases staggered like bench.py), then a few profiled steps.
  ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:k_step -c 2 \
      -o gpurun_out/k_step_r02 python scripts/gpu_ncu_target.py
(the 500 pre-roll steps run unprofiled; cudaProfilerStart brackets the captured launches)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench

class A: pass
args = A(); args.benchmark = sys.argv[1] if len(sys.argv) > 1 else "MT50"; args.envs_per_gpu = 4096; args.seed = 42
env, names, n_full, kind = bench.build_env(args, 0, 0)
env.reset()
bench.stagger(env, args.seed)
g = torch.Generator(device=env.device); g.manual_seed(1)
for i in range(500):
    env.step_torch(torch.rand(4096, 4, device=env.device, generator=g) * 2 - 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(3):
    env.step_torch(torch.rand(4096, 4, device=env.device, generator=g) * 2 - 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", env.engine.counters())
