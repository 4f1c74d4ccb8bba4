A="--steps 30 --warmup 3 --e2e-steps 3 --cpu-steps-per-env 1"
(MW_BENCH_NOFLUSH=1 MW_BENCH_FAKE_RANK=1 timeout 60 python bench.py $A > /dev/null 2> /tmp/r1.err; grep '\[bench\]' /tmp/r1.err | cut -c1-150) &
wait
