mkdir -p gpurun_out
export MW_B200_LIB=$PWD/tests/_build/libmwb200_next.so
python scripts/gpu_ab.py gpurun_out/ab_next3.json > gpurun_out/ab_next3.log 2>&1
for h in 0 -1 4 3 2; do
MW_B200_HEAD_WARPS=$h python bench.py --steps 100 --warmup 5 > gpurun_out/bench21_h$h.json 2> gpurun_out/bench21_h$h.err
done
python - <<'PY'
import json
for n in ("0", "-1", "4", "3", "2"):
    try:
        d = json.loads(open(f"gpurun_out/bench21_h{n}.json").read().strip().split("\n")[-1])
        print("head", n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench21_h{n}.err").read()[-1500:])
PY
python scripts/gpu_host_breakdown.py > gpurun_out/host_breakdown.txt 2>&1; tail -3 gpurun_out/host_breakdown.txt
python scripts/gpu_cost_dist.py > gpurun_out/cost_dist21.txt 2>&1; grep -A2 "^step 300\|^step 400" gpurun_out/cost_dist21.txt
