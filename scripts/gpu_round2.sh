#!/bin/bash
# Round-2 batched GPU session at HEAD: parity tests, smoke, bench lines of every BASELINE config, per-task table, cost
# distribution, ncu launch list of the timed region + one full capture of k_step in steady state.  Outputs -> gpurun_out/,
# summarised into profiles/r02_* by scripts/make_profiles.py r02.
mkdir -p gpurun_out; rm -f gpurun_out/contact_rich.csv gpurun_out/open_loop.csv gpurun_out/long_rollout.csv gpurun_out/policy_success.csv
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rx -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_mt50.json 2> gpurun_out/bench_mt50.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_mt50.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_mt50_driver_window.json 2> /dev/null; cut -c1-160 gpurun_out/bench_mt50_driver_window.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 python bench.py --benchmark MT10 > gpurun_out/bench_mt10.json 2> gpurun_out/bench_mt10.err; cut -c1-160 gpurun_out/bench_mt10.json
timeout 600 python bench.py --benchmark reach-v3 > gpurun_out/bench_reach.json 2> gpurun_out/bench_reach.err; cut -c1-160 gpurun_out/bench_reach.json
timeout 900 python bench.py --benchmark ML45-train --envs-per-gpu 8192 > gpurun_out/bench_ml45_train.json 2> gpurun_out/bench_ml45_train.err; cut -c1-160 gpurun_out/bench_ml45_train.json
timeout 900 python bench.py --benchmark ML45-test --envs-per-gpu 8192 > gpurun_out/bench_ml45_test.json 2> gpurun_out/bench_ml45_test.err; cut -c1-160 gpurun_out/bench_ml45_test.json
timeout 600 python scripts/gpu_task_times.py > gpurun_out/task_times.jsonl 2> gpurun_out/task_times.err; echo "task times rc=$?"
timeout 600 python scripts/gpu_cost_dist.py > gpurun_out/cost_distribution.txt 2>&1; tail -3 gpurun_out/cost_distribution.txt
timeout 300 python scripts/gpu_host_breakdown.py > gpurun_out/host_breakdown.txt 2>&1; tail -3 gpurun_out/host_breakdown.txt
MW_BENCH_NCU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 60 --warmup 5 --e2e-steps 3 --cpu-steps-per-env 5 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_step -c 1 -f -o gpurun_out/k_step_full python scripts/gpu_ncu_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -40
