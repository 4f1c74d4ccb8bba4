mkdir -p gpurun_out
P=$PWD/metaworld_b200
for v in vc vd; do
MW_B200_LIB=$P/libmwb200_$v.so python scripts/gpu_ab.py gpurun_out/ab_$v.json > gpurun_out/ab_$v.log 2>&1 || tail -5 gpurun_out/ab_$v.log
done
echo "== vc vs vd"; python scripts/gpu_ab.py --cmp gpurun_out/ab_vc.json gpurun_out/ab_vd.json | tail -1
for v in vd vc; do
MW_B200_LIB=$P/libmwb200_$v.so timeout 600 python bench.py --steps 100 --warmup 5 --cpu-steps-per-env 5 --e2e-steps 20 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$v.json").read().strip().split("\n")[-1])
    print("$v", round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], {k: d["phases"][k] for k in ("kin_mass","solver","bias_smooth","euler_glue","warp_cycles_per_env_step_own_work")})
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_$v.err").read()[-1500:])
PY
done
