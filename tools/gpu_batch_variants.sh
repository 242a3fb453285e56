#!/bin/bash
# 100-permutation points again, now that profiles/r2_counts.json holds the instruction counts of the final split-carrier kernels
set -u
OUT=gpurun_out
python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_config3_p100.json 2> $OUT/r2_bench_config3_p100.err
python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 100 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2_config5_n1_p100.json 2> $OUT/r2_config5_n1_p100.err
tail -n 1 $OUT/r2_bench_config3_p100.err $OUT/r2_config5_n1_p100.err
