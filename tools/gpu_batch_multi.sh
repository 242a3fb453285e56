#!/bin/bash
# Round-2 multi-GPU batch: bash tools/gpu_batch_multi.sh N   (run through `gpurun --gpus N` from the repo root)
set -u
N=$1
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nproc > $OUT/r2_box_n$N.txt; free -g | head -2 >> $OUT/r2_box_n$N.txt; nvidia-smi -L >> $OUT/r2_box_n$N.txt
if [ "$N" = "2" ]; then
  python -m pytest tests -m gpu -q > $OUT/r2_gputest_n2.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_n2.log
fi
# BASELINE config 3 on N GPUs: one block of 1,000 permutations per GPU (weak), plus the fixed job row-sharded (`strong` object) and both parity checks
$TR bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "rc=$?" >> $OUT/r2_bench_n$N.err
# BASELINE config 4: 100,000 patients, path length 5, 10,000 permutations in total over the N GPUs (value table generated on the device)
P=$((10000 / N))
$TR bench.py --gpus $N --n-cases 50000 --n-ctrls 50000 --path-length 5 --n-perms $P --steps 2 --warmup 3 --no-e2e --no-strong > $OUT/r2_cfg4_n$N.json 2> $OUT/r2_cfg4_n$N.err; echo "rc=$?" >> $OUT/r2_cfg4_n$N.err
if [ "$N" = "8" ]; then
  # BASELINE config 5 on 8 GPUs: 50,000 patients, path length 4; 8,000 and 100,000 permutations in total
  for P in 1000 12500; do
    $TR bench.py --gpus $N --n-cases 25000 --n-ctrls 25000 --n-perms $P --steps 2 --warmup 3 --no-e2e --no-strong > $OUT/r2_cfg5_n8_p$P.json 2> $OUT/r2_cfg5_n8_p$P.err; echo "rc=$?" >> $OUT/r2_cfg5_n8_p$P.err
  done
fi
for f in $OUT/r2_bench_n$N.err $OUT/r2_cfg4_n$N.err; do tail -n 2 $f; done
