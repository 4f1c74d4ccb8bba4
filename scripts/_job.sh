mkdir -p gpurun_out
python -m pytest tests/test_gpu.py -q -x -k "wrapped_single or (open_loop and drawer)" 2>&1 | tail -3
python scripts/gpu_ab.py gpurun_out/ab_off.json > gpurun_out/ab_off.log 2>&1
MW_B200_SPLIT_FRAC=0.65 python scripts/gpu_ab.py gpurun_out/ab_split.json > gpurun_out/ab_split.log 2>&1
python scripts/gpu_ab.py --cmp gpurun_out/ab_off.json gpurun_out/ab_split.json
for v in 0:4 0.5:4 0.65:4 0.8:4 0.65:3 0.65:5; do
f=${v%%:*}; w=${v##*:}
MW_B200_SPLIT_FRAC=$f MW_B200_SPLIT_WARPS=$w python bench.py --steps 100 --warmup 5 --cpu-steps-per-env 20 --e2e-steps 30 > gpurun_out/bench30_$f-$w.json 2> gpurun_out/bench30_$f-$w.err
done
python - <<'PY'
import json
for n in ("0-4", "0.5-4", "0.65-4", "0.8-4", "0.65-3", "0.65-5"):
    try:
        d = json.loads(open(f"gpurun_out/bench30_{n}.json").read().strip().split("\n")[-1])
        print("split", n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench30_{n}.err").read()[-1500:])
PY
