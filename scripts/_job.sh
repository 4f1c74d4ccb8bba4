mkdir -p gpurun_out
P=$PWD/metaworld_b200
MW_B200_LIB=$P/libmwb200_prev.so python scripts/gpu_ab.py gpurun_out/ab_prev.json > gpurun_out/ab_prev.log 2>&1 || tail -5 gpurun_out/ab_prev.log
python scripts/gpu_ab.py gpurun_out/ab_new.json > gpurun_out/ab_new.log 2>&1 || tail -5 gpurun_out/ab_new.log
python scripts/gpu_ab.py --cmp gpurun_out/ab_prev.json gpurun_out/ab_new.json
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_step -c 1 -f -o gpurun_out/k_step_full python scripts/gpu_ncu_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head
