mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 100 python -m pytest tests/test_gpu.py -q -x -p no:cacheprovider -k "heterogeneous or determinis or mt10_full_size or resume" 2>&1 | tail -3
