#!/bin/bash
# One batched GPU session: parity tests, bench (both arms), ncu launch list + full capture of the step kernel.
mkdir -p gpurun_out; rm -f gpurun_out/contact_rich.csv
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -3
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_mt50.json 2> gpurun_out/bench_mt50.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_mt50.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 3 --cpu-steps-per-env 5 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 4 -c 1 -o gpurun_out/k_step_full -f python bench.py --steps 2 --warmup 3 --e2e-steps 3 --cpu-steps-per-env 5 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
