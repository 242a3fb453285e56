#!/bin/bash
# BASELINE config 5 on 8 GPUs with the final build: 50,000 patients, path length 4, 1,000 permutations per GPU (8,000 in total)
set -u
N=8
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --n-cases 25000 --n-ctrls 25000 --n-perms 1000 --steps 2 --warmup 3 --no-e2e --no-strong > $OUT/r2_cfg5_n8_p1000.json 2> $OUT/r2_cfg5_n8_p1000.err; echo "rc=$?" >> $OUT/r2_cfg5_n8_p1000.err
tail -n 2 $OUT/r2_cfg5_n8_p1000.err
