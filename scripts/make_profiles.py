"""Builds the tracked summaries under profiles/ from the scratch outputs of scripts/gpu_round.sh (gpurun_out/)."""
import csv, json, os, subprocess, sys, collections
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)

def jl(path): return json.loads(open(path).read().strip().split("\n")[-1])     # (a library banner may precede the JSON line)

# 1. bench lines (one per BASELINE config that fits one GPU, + both arms of the headline, + the driver's 20-step window)
for src, dst in (("bench_mt50.json", f"{R}_bench_mt50.json"), ("bench_ref.json", f"{R}_bench_reference.json"),
                 ("bench_mt50_driver_window.json", f"{R}_bench_mt50_driver_window.json"), ("bench_mt10.json", f"{R}_bench_mt10.json"),
                 ("bench_reach.json", f"{R}_bench_reach_v3.json"), ("bench_ml45_train.json", f"{R}_bench_ml45_train.json"),
                 ("bench_ml45_test.json", f"{R}_bench_ml45_test.json"), ("bench_2gpu_gather.json", f"{R}_bench_2gpu_gather.json"),
                 ("bench_2gpu.json", f"{R}_bench_2gpu.json"), ("bench_8gpu.json", f"{R}_bench_8gpu.json"), ("bench_8gpu_gather.json", f"{R}_bench_8gpu_gather.json")):
    if os.path.exists(os.path.join(G, src)):
        try:
            d = jl(os.path.join(G, src)); json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
        except Exception as e:
            print("skip", src, e)

# 2. launch list -> per-kernel summary (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: shares only)
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if r and r[0].isdigit()]
hdr = next(r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if r and r[0] == "ID")
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    k = r[iK].split("(")[0]; v = float(r[iV].replace(",", "")); u = r[iU]
    v_us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v_us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{R}_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,total_us,mean_us,share_of_gpu_time\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{t:.1f},{t/n:.1f},{t/tot:.4f}\n")

# 3. full capture of k_step: key metrics + source hot spots
rep = os.path.join(G, "k_step_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u, v = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__cycles_elapsed.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for i, hname in enumerate(h):
    if hname.startswith("smsp__pcsamp_warps_issue_stalled_") and not hname.endswith("_not_issued"): want.append(hname)
with open(os.path.join(P, f"{R}_k_step_ncu_summary.csv"), "w") as f:
    f.write("metric,value,unit\n")
    for w in want:
        if w in h: i = h.index(w); f.write(f"{w},{v[i]},{u[i]}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
hot = subprocess.run([sys.executable, "scripts/ncu_lines.py", "/tmp/_src.csv", "30"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{R}_k_step_hotspots.txt"), "w").write("k_step, MT50 4096 envs, steady state (episode phases uniform over 0..499, scripts/gpu_ncu_target.py; ncu --set full --import-source on; stall samples and executed warp instructions by function / line)\n" + hot)
if os.path.exists(os.path.join(G, "cost_distribution.txt")):
    open(os.path.join(P, f"{R}_cost_distribution.txt"), "w").write(open(os.path.join(G, "cost_distribution.txt")).read())
if os.path.exists(os.path.join(G, "host_breakdown.txt")):
    open(os.path.join(P, f"{R}_host_breakdown.txt"), "w").write("MetaWorldVecEnv.step (numpy API), MT50 @ 4096, steady state; scripts/gpu_host_breakdown.py\n" + open(os.path.join(G, "host_breakdown.txt")).read())

# 4. per-task table
rows = [json.loads(l) for l in open(os.path.join(G, "task_times.jsonl"))]
rows.sort(key=lambda r: -r["mcycles"])
with open(os.path.join(P, f"{R}_task_times.md"), "w") as f:
    f.write("Per-task step cost: 888 envs of one task, 20 random-action steps after 3 warm-up steps (scripts/gpu_task_times.py).\n"
            "`Mcyc` = mean OWN-WORK warp cycles per env step (1 warp = 1 env; time spent waiting for CTA-mates at phase barriers excluded); phases as fractions of that; events per env step.\n\n")
    f.write("| task | ms/step | Mcyc | collide | of which GJK/EPA | solver | convex pairs | EPA expansions | GJK iters | Newton its/pass |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r['task']} | {r['ms']:.2f} | {r['mcycles']:.2f} | {r['collide']:.2f} | {r['gjk']:.2f} | {r['solver']:.2f} | {r['pairs']:.1f} | {r['epa']:.1f} | {r['gjkit']:.1f} | {r['newton']:.2f} |\n")

# 5. parity table
ol = {r[0]: r for r in csv.reader(open(os.path.join(G, "open_loop.csv")))}
cr = {r[0]: r for r in csv.reader(open(os.path.join(G, "contact_rich.csv")))}
xf = [l.strip() for l in open(os.path.join(G, "pytest.log")) if l.startswith("XFAIL")]
with open(os.path.join(P, f"{R}_parity.md"), "w") as f:
    f.write("Device (float32 step, float64 collision, float64 reset snapshots) vs goldens (the reference's env classes run unmodified on restated float64 physics), through the C ABI on a B200; obs, reward and all 7 info keys.\n"
            "Open loop: 60 random-action steps from reset, 3 goals, worst absolute error over the rollout.  Contact rich: single steps teacher-forced\n"
            "from the oracle's state along trajectories driven by the reference's scripted policy (238 steps per task).\n\n")
    f.write("| task | open-loop worst obs err | open-loop worst reward err | contact-rich: frac of steps within 1e-4 | median err | p90 err | worst err |\n|---|---|---|---|---|---|---|\n")
    for t in sorted(set(ol) | set(cr)):
        o = ol.get(t, ["", "", "-", "-"]); c = cr.get(t, ["", "", "-", "-", "-", "-"])
        f.write(f"| {t} | {o[2]} | {o[3]} | {c[2]} | {c[3]} | {c[4]} | {c[5]} |\n")
    lr = {r[0]: r for r in csv.reader(open(os.path.join(G, "long_rollout.csv")))} if os.path.exists(os.path.join(G, "long_rollout.csv")) else {}
    ps = {r[0]: r for r in csv.reader(open(os.path.join(G, "policy_success.csv")))} if os.path.exists(os.path.join(G, "policy_success.csv")) else {}
    f.write("\nFull 500-step open-loop episodes (worst |obs, reward, info| error up to step 60 / 125 / 250 / 500) and scripted-policy action replays "
            "(successes out of 5 goals: device / reference glue on the oracle):\n\n| task | <=60 | <=125 | <=250 | <=500 | policy replay device | reference glue |\n|---|---|---|---|---|---|---|\n")
    for t in sorted(set(lr) | set(ps)):
        a = lr.get(t, [t, "-", "-", "-", "-"]); b = ps.get(t, [t, "-", "-"])
        f.write(f"| {t} | {a[1]} | {a[2]} | {a[3]} | {a[4]} | {b[1]} | {b[2]} |\n")
    summ = [l.strip() for l in open(os.path.join(G, "pytest.log")) if " passed" in l or " failed" in l]
    f.write("\nGPU suite of this run: " + "; ".join(summ) + "\n")
    f.write("\nxfail list of this run (pytest -rx):\n\n")
    for l in xf: f.write(f"* `{l}`\n")
print(open(os.path.join(P, f"{R}_launches_summary.csv")).read())
print(open(os.path.join(P, f"{R}_k_step_ncu_summary.csv")).read())

# ---- static SASS summary of the shipped library, per kernel (cuobjdump; no GPU needed)
import re, subprocess
so = os.path.join(os.path.dirname(P), "metaworld_b200", "libmwb200.so")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
    usage = dict(re.findall(r"Function (\S+):\n\s+(REG:\d+ STACK:\d+ SHARED:\d+)", res))
    parts = re.split(r"\n\s+Function : (\S+)\n", sass)
    keys = ["FFMA", "DFMA", "STL", "LDL", "LDS", "STS", "SHFL", "WARPSYNC", "BAR.SYNC", "MUFU", "UBLKCP", "SYNCS"]
    with open(os.path.join(P, f"{R}_sass_summary.md"), "w") as f:
        f.write("# Static SASS instruction counts per kernel of `metaworld_b200/libmwb200.so` (sm_100a; `cuobjdump -sass`)\n\n")
        f.write("`UBLKCP` + `SYNCS` = TMA bulk copy of the model blob + mbarrier.  `k_step` is the kernel the bench times; `k_snapshot` "
                "(reset double pass), `k_substeps` (diagnostics) and `k_evaluate` (single-env `evaluate_state`) instantiate the same physics.\n\n")
        f.write("| kernel | instructions | " + " | ".join(keys) + " | resources |\n|---|---|" + "---|" * (len(keys) + 1) + "\n")
        for name, body in zip(parts[1::2], parts[2::2]):
            n = len(re.findall(r"/\*[0-9a-f]{4,}\*/", body))
            mm = re.match(r"_Z(\d+)", name); short = name[len(mm.group(0)):][: int(mm.group(1))] if mm else name
            f.write(f"| `{short}` | {n} | " + " | ".join(str(len(re.findall(r"\b" + re.escape(k), body))) for k in keys) + f" | {usage.get(name, '')} |\n")
    print(open(os.path.join(P, f"{R}_sass_summary.md")).read())
