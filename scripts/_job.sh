mkdir -p gpurun_out
python scripts/gpu_diag.py open drawer-close-v3 > gpurun_out/diag_dc64.log 2>&1
MW_DIAG_RV32=1 python scripts/gpu_diag.py open drawer-close-v3 > gpurun_out/diag_dc32.log 2>&1
python -m pytest tests/test_gpu.py -q -x -k "wrapped_single or (open_loop and drawer)" 2>&1 | tail -5
tail -25 gpurun_out/diag_dc64.log; tail -8 gpurun_out/diag_dc32.log
