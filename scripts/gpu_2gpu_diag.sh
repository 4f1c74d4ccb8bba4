#!/bin/bash
# Why is rank 1 slower under torchrun than the same GPU alone?  (a) device index vs (b) rank-derived random streams.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
A="--steps 40 --warmup 5 --e2e-steps 5 --cpu-steps-per-env 1"
(MW_BENCH_FAKE_RANK=1 CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py $A > /dev/null 2> gpurun_out/diag_fake1.err &
 MW_BENCH_FAKE_RANK=0 CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py $A > /dev/null 2> gpurun_out/diag_fake0.err &
 wait)
echo "solo, rank-1 streams on GPU 0:"; grep "\[bench\]" gpurun_out/diag_fake1.err
echo "solo, rank-0 streams on GPU 1:"; grep "\[bench\]" gpurun_out/diag_fake0.err
MW_BENCH_SWAP=1 MW_BENCH_PG=gloo timeout 300 $T bench.py --gpus 2 $A > gpurun_out/diag_swap.json 2> gpurun_out/diag_swap.err; echo "torchrun, devices swapped:"; grep "\[bench\]" gpurun_out/diag_swap.err
