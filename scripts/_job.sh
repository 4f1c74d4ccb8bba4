mkdir -p gpurun_out
MW_B200_LIB=$PWD/tests/_build/libmwb200_next.so python scripts/gpu_ab.py gpurun_out/ab_next.json > gpurun_out/ab_next.log 2>&1
python scripts/gpu_ab.py --cmp gpurun_out/ab_cur.json gpurun_out/ab_next.json | tail -2
MW_B200_HOST_OUTPUTS=1 python bench.py --steps 100 --warmup 5 > gpurun_out/bench18_ho.json 2> gpurun_out/bench18_ho.err
MW_B200_LIB=$PWD/tests/_build/libmwb200_next.so python bench.py --steps 100 --warmup 5 > gpurun_out/bench18_next.json 2> gpurun_out/bench18_next.err
python - <<'PY'
import json
for n in ("ho", "next"):
    try:
        d = json.loads(open(f"gpurun_out/bench18_{n}.json").read().strip().split("\n")[-1])
        print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["clocks"], d["config"]["build"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench18_{n}.err").read()[-1500:])
PY
python scripts/gpu_host_breakdown.py > gpurun_out/host_breakdown.txt 2>&1; tail -4 gpurun_out/host_breakdown.txt
MW_B200_HOST_OUTPUTS=1 python scripts/gpu_host_breakdown.py > gpurun_out/host_breakdown_ho.txt 2>&1; tail -4 gpurun_out/host_breakdown_ho.txt
python scripts/gpu_cost_dist.py > gpurun_out/cost_dist18.txt 2>&1; grep -A1 "^step 300\|^step 400" gpurun_out/cost_dist18.txt
