#!/bin/bash
# the default bench line once more, now that profiles/r2_counts.json holds the instruction counts of the final kernels
set -u
OUT=gpurun_out
python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_n1.json 2> $OUT/r2_bench_n1.err; echo "rc=$?" >> $OUT/r2_bench_n1.err
tail -n 2 $OUT/r2_bench_n1.err
