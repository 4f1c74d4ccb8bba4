#!/bin/bash
# 2-GPU weak-scaling lines (gpurun --gpus 2): the step path has no collective; --gather adds the optional NCCL epilogue of
# BASELINE config 4 (obs + packed reward/info/flags of every rank, side stream).  Outputs -> gpurun_out/bench_2gpu*.json
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --steps 100 --warmup 5 --e2e-steps 30 --cpu-steps-per-env 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"; cut -c1-200 gpurun_out/bench_2gpu.json
timeout 600 $T bench.py --gpus 2 --gather --steps 100 --warmup 5 --e2e-steps 30 --cpu-steps-per-env 5 > gpurun_out/bench_2gpu_gather.json 2> gpurun_out/bench_2gpu_gather.err; echo "2gpu gather rc=$?"; cut -c1-200 gpurun_out/bench_2gpu_gather.json
grep -h "NCCL INFO.*\(NVLS\|Connected\|nranks\)" gpurun_out/bench_2gpu_gather.err | head -5
