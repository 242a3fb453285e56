#!/bin/bash
# BASELINE config 4 with the final build: bash tools/gpu_batch_multi4.sh N  (through `gpurun --gpus N`)
# 100,000 patients, path length 5, 10,000 permutations in total over the N GPUs (value table generated on the device)
set -u
N=$1
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
P=$((10000 / N))
$TR bench.py --gpus $N --n-cases 50000 --n-ctrls 50000 --path-length 5 --n-perms $P --steps 2 --warmup 3 --no-e2e --no-strong > $OUT/r2_cfg4_n$N.json 2> $OUT/r2_cfg4_n$N.err; echo "rc=$?" >> $OUT/r2_cfg4_n$N.err
tail -n 2 $OUT/r2_cfg4_n$N.err
