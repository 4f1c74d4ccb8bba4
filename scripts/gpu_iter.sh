#!/bin/bash
# Iteration run: parity tests + diagnostics for a list of tasks + short bench.
mkdir -p gpurun_out; rm -f gpurun_out/contact_rich.csv
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^FAILED|passed|failed" gpurun_out/pytest.log | tail -40
python scripts/gpu_diag.py reset box-close-v3 disassemble-v3 peg-unplug-side-v3 reach-wall-v3 > gpurun_out/diag_reset.txt 2>&1
python scripts/gpu_diag.py open coffee-pull-v3 hammer-v3 handle-press-v3 reach-wall-v3 peg-unplug-side-v3 > gpurun_out/diag_open.txt 2>&1
python scripts/gpu_diag.py cr box-close-v3 window-close-v3 button-press-wall-v3 plate-slide-side-v3 soccer-v3 button-press-v3 bin-picking-v3 > gpurun_out/diag_cr.txt 2>&1
timeout 600 python bench.py --steps 60 --warmup 5 --e2e-steps 10 --cpu-steps-per-env 20 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cut -c1-330 gpurun_out/bench_quick.json
