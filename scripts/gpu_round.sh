#!/bin/bash
# One batched GPU session: parity tests, smoke, bench (both arms), per-task table, ncu launch list + full capture of k_step.
mkdir -p gpurun_out; rm -f gpurun_out/contact_rich.csv gpurun_out/open_loop.csv
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -rx -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
grep -E "passed|failed" gpurun_out/pytest.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_mt50.json 2> gpurun_out/bench_mt50.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_mt50.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-200 gpurun_out/bench_ref.json
timeout 600 python scripts/gpu_task_times.py > gpurun_out/task_times.jsonl 2> gpurun_out/task_times.err; echo "task times rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 60 --warmup 3 --e2e-steps 3 --cpu-steps-per-env 5 > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 100 -c 1 -o gpurun_out/k_step_full -f python bench.py --steps 110 --warmup 3 --e2e-steps 3 --cpu-steps-per-env 5 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | head -30
