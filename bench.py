#!/usr/bin/env python
"""Benchmark: env steps/s of the batched Meta-World step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU implementation of the same
                                                            # path on the host cores (see below), rank 0 only

A "step" is one VectorEnv.step over all environments of a rank (4096 by default): per env 5 physics substeps
+ 1 forward pass + obs + reward + autoreset.  `value` is measured with actions and state resident in HBM
(CUDA events around each mw_step launch, max over ranks); `e2e` goes through the public numpy API
(`MetaWorldVecEnv.step`) with host actions in and host obs/reward/flags/info out every step.

Reference arm: MuJoCo / gymnasium (the reference's physics + glue) are not installed in this image and cannot
be installed offline, so `oracle/` -- the float64 CPU restatement of the same path -- is what is timed, on all
host cores (one process per core, each stepping its own envs), labelled kind="port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env steps/sec MT50 batched (4096 envs/GPU) vs CPU MuJoCo"


def implemented_tasks(benchmark):
    from metaworld_b200 import benchmarks as B
    from metaworld_b200.tasks import TASKS
    names = {"MT50": B.MT50, "MT10": B.MT10, "MT25": B.MT25}.get(benchmark, [benchmark])
    return [n for n in names if n in TASKS], len(names)


def algorithmic_bytes(names):
    """per env step 4*(2*nq + 4*nv + 120) bytes (fp32 state, 5 substeps fused)."""
    from metaworld_b200 import modelzoo
    from metaworld_b200.tasks import TASKS
    b = []
    for n in names:
        m = modelzoo.full_model(TASKS[n].xml)
        b.append(4 * (2 * m.nq + 4 * m.nv + 120))
    return float(np.mean(b))


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


# ----------------------------------------------------------------------------- CPU (oracle) timing
def _cpu_worker(args):
    names, steps, seed = args
    from oracle.tasks import TASKS as OT
    from metaworld_b200 import benchmarks as B
    envs = []
    for i, n in enumerate(names):
        e = OT[n]()
        t = B.make_tasks([n], False, seed=seed + i, n_goals=1)[0].unpack()
        e.set_task_vec(t["rand_vec"], False)
        e.reset()
        envs.append(e)
    rng = np.random.default_rng(seed)
    t0 = time.perf_counter()
    n = 0
    for s in range(steps):
        for e in envs:
            if e.curr_path_length >= e.max_path_length:
                e.reset()
            e.step(rng.uniform(-1, 1, 4).astype(np.float32))
            n += 1
    return n, time.perf_counter() - t0


def usable_cores():
    """Host threads this process may actually use: the affinity mask capped by the cgroup CPU quota (the GPU boxes
    show 128 logical cores but cpu.max = 16 CPUs)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        pass
    return max(1, n)


def cpu_baseline(names, cores, steps_per_env, seed=42):
    """Times the CPU restatement (oracle) of the same path: `cores` processes, each with its own sub-envs; the task
    types are dealt round-robin over the processes so every type of the workload is in the sample."""
    import multiprocessing as mp
    per = max(1, -(-len(names) // cores))
    jobs = [([names[(c * per + k) % len(names)] for k in range(per)], steps_per_env, seed + 1000 * c) for c in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as p:
            res = p.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, f"{cores} process(es) x {per} envs x {steps_per_env} random-action steps (float64 CPU restatement, wall {wall:.1f}s)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import mjphys
    mjphys.build()
    names, _ = implemented_tasks(args.benchmark)
    cores = usable_cores()
    vals = []
    sample = ""
    t_all = time.perf_counter()
    for _ in range(args.warmup + args.steps):
        v, sample = cpu_baseline(names, cores, args.ref_steps_per_env)
        vals.append(v)
        if time.perf_counter() - t_all > 150:
            break
    vals = vals[min(args.warmup, len(vals) - 1):]
    value = float(np.mean(vals))
    per_step_envs = args.envs_per_gpu * args.gpus
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env_steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_step_envs / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.benchmark}: {len(names)} task types, CPU restatement of the same step path, random actions U(-1,1)",
                       "tasks": names, "note": "reference CPU MuJoCo could not be executed (mujoco/gymnasium not installed, no network); timed: oracle/ float64 restatement"},
            "cpu_baseline": {"value": value, "unit": "env_steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from metaworld_b200.vector_env import MetaWorldVecEnv
    from metaworld_b200 import benchmarks as B
    from metaworld_b200.engine import lib

    names, n_full = implemented_tasks(args.benchmark)
    N = args.envs_per_gpu
    # every rank owns all task types (balanced shards, no collective on the step path); goals differ per rank via the seed
    tasks_all = B.make_tasks(names, False, seed=args.seed)
    tasks = [[t for t in tasks_all if t.env_name == n] for n in names]
    env = MetaWorldVecEnv(names, tasks, num_envs=N, seed=args.seed + 1000 * rank, use_one_hot=True, num_tasks=max(n_full, len(names)),
                          max_episode_steps=500, device=local)
    env.reset()
    env.enable_device_sampler()
    dev = env.device
    K, W = args.steps, args.warmup
    gen = torch.Generator(device=dev); gen.manual_seed(args.seed + rank)
    actions = torch.rand(K + W, N, 4, device=dev, generator=gen) * 2 - 1      # resident in HBM before timing
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)                    # > 126 MB L2
    sampler = ClockSampler(local); sampler.start()
    for i in range(W):
        env.step_torch(actions[i])
    torch.cuda.synchronize()
    env.engine.profile()                            # clear the per-phase cycle counters
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    torch.cuda.synchronize()
    for i in range(K):
        flush.fill_(float(i))                       # evict L2 between timed iterations (outside the event pair)
        ev[i][0].record()
        env.step_torch(actions[W + i])
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    prof = env.engine.profile()
    # ---- e2e through the numpy API: pinned H2D of actions, D2H of obs/reward/flags/info inside the timed region
    Ke = max(3, min(K, args.e2e_steps))
    a_host = (np.random.default_rng(args.seed + rank).uniform(-1, 1, size=(Ke + 2, N, 4))).astype(np.float32)
    env._device_sampler = False
    for i in range(2):
        env.step(a_host[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        env.step(a_host[2 + i])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    counters = env.engine.counters()
    if rank == 0:
        value = N * world * K / (ms * 1e-3)
        e2e_val = N * world * Ke / (e2e_ms * 1e-3)
        bytes_step = algorithmic_bytes(names)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = (value / world) * bytes_step / 1e9
        traffic = None      # dram bytes per k_step launch from the committed `ncu --set full` capture of this round
        try:
            import csv as _csv, glob as _glob
            prof_csv = sorted(_glob.glob(os.path.join(ROOT, "profiles", "r*_k_step_ncu_summary.csv")))[-1]
            m = {r[0]: (float(r[1]), r[2]) for r in _csv.reader(open(prof_csv)) if len(r) == 3 and r[0].startswith("dram__bytes")}
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = sum(v * scale.get(u, 1.0) for v, u in m.values()) if m else None
        except Exception:
            traffic = None
        cpu_val, cpu_sample = cpu_baseline(names, 1, args.cpu_steps_per_env)
        obs_dim = env.obs_dim
        line = {"metric": METRIC, "value": value, "unit": "env_steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.benchmark}: {len(names)} of {n_full} task types implemented, {N} envs/GPU interleaved by task id, "
                                       "random actions U(-1,1), 500-step episodes with SAME_STEP autoreset, one-hot obs",
                           "tasks": names, "envs_per_gpu": N, "l2": "256 MB device write between timed steps (outside the per-step CUDA-event pairs)",
                           "build": lib().mw_build_info().decode(), "sharding": "env-parallel, no collective on the step path"},
                "e2e": {"value": e2e_val, "unit": "env_steps/s", "h2d_bytes_per_step": int(N * 4 * 4 + N * 4),
                        "d2h_bytes_per_step": int(N * (obs_dim + 9) * 4), "steps": Ke},
                "gpu_launches": 3 * (K + W + Ke + 2),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "algorithmic_bytes_per_launch": bytes_step * N, "algorithmic_bytes_per_env_step": bytes_step,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                             "note": "latency bound: one warp runs one env's dependent chain (nv<=17, <1 KB state per env step) and the step ends with the slowest env; "
                                     "measured dram traffic above the algorithmic bytes is per-thread stack (local memory) spilling past L2, not re-reads of state; see DESIGN.md section 6"},
                "cpu_baseline": {"value": cpu_val, "unit": "env_steps/s", "cores": 1, "kind": "port", "sample": cpu_sample},
                "clocks": sampler.summary(),
                "phases": {"unit": "fraction of per-warp step cycles (clock64), timed region",
                           **{k: round(prof[k] / max(1, prof["step"]), 4) for k in list(prof)[:8]},
                           "warp_cycles_per_env_step": prof["step"] / max(1, N * K),
                           "convex_pairs_per_env_step": prof["n_convex_pairs"] / max(1, N * K),
                           "epa_expansions_per_env_step": prof["n_epa_expansions"] / max(1, N * K),
                           "gjk_iters_per_env_step": prof["n_gjk_iters"] / max(1, N * K)},
                "solver": {"mean_newton_iters_per_pass": counters["solver_iters"] / max(1, counters["forward_passes"]),
                           "contacts_dropped": counters["contacts_dropped"]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)      # one full 500-step episode incl. autoreset
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--benchmark", default="MT50")
    ap.add_argument("--envs-per-gpu", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--e2e-steps", type=int, default=500)
    ap.add_argument("--cpu-steps-per-env", type=int, default=600)
    ap.add_argument("--ref-steps-per-env", type=int, default=1000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
