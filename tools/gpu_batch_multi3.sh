#!/bin/bash
# Multi-GPU re-measurement of BASELINE config 3 with the final build: bash tools/gpu_batch_multi3.sh N  (through `gpurun --gpus N`)
# (weak: one block of 1,000 permutations per GPU; `strong`: the fixed job row-sharded; e2e; both parity checks)
set -u
N=$1
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "rc=$?" >> $OUT/r2_bench_n$N.err
tail -n 2 $OUT/r2_bench_n$N.err
