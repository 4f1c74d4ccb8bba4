mkdir -p gpurun_out
python scripts/gpu_ab.py gpurun_out/ab_cur.json > gpurun_out/ab_cur.log 2>&1
MW_B200_LIB=$PWD/tests/_build/libmwb200_w8.so python scripts/gpu_ab.py gpurun_out/ab_w8.json > gpurun_out/ab_w8.log 2>&1
python scripts/gpu_ab.py --cmp gpurun_out/ab_cur.json gpurun_out/ab_w8.json | tail -2
python bench.py --steps 100 --warmup 5 > gpurun_out/bench18_cur.json 2> gpurun_out/bench18_cur.err
MW_B200_LIB=$PWD/tests/_build/libmwb200_w8.so python bench.py --steps 100 --warmup 5 > gpurun_out/bench18_w8.json 2> gpurun_out/bench18_w8.err
python - <<'PY'
import json
for n in ("cur", "w8"):
    try:
        d = json.loads(open(f"gpurun_out/bench18_{n}.json").read().strip().split("\n")[-1])
        print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["clocks"], d["config"]["build"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench18_{n}.err").read()[-1500:])
PY
python scripts/gpu_host_breakdown.py > gpurun_out/host_breakdown.txt 2>&1; head -45 gpurun_out/host_breakdown.txt; tail -3 gpurun_out/host_breakdown.txt
