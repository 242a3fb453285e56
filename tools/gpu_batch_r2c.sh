#!/bin/bash
# Round-2 batch 3 (one B200): A/B of the thresholded look-up variants with the exact-loop fix, each with parity; CPU sub-shape baseline at 50,000 patients
set -u
OUT=gpurun_out
for V in v2b v3b; do
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python -m pytest tests/test_join_gpu.py tests/test_fullsize_gpu.py -q -x -k "schedule_matches_oracle or golden or many_permutations or dense_carrier or fullsize or agree or compose or first_rows" > $OUT/r2_gputest_$V.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_$V.log
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python bench.py --steps 10 --warmup 3 --e2e-steps 10 > $OUT/r2_bench_$V.json 2> $OUT/r2_bench_$V.err; echo "rc=$?" >> $OUT/r2_bench_$V.err
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_${V}_p100.json 2> $OUT/r2_bench_${V}_p100.err
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python bench.py --path-length 5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_${V}_len5.json 2> $OUT/r2_bench_${V}_len5.err
done
python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 1000 --steps 3 --warmup 3 --no-e2e --cpu-seconds 8 > $OUT/r2_cfg5_n1_p1000_cpu.json 2> $OUT/r2_cfg5_n1_p1000_cpu.err; echo "rc=$?" >> $OUT/r2_cfg5_n1_p1000_cpu.err
for f in $OUT/r2_gputest_v2b.log $OUT/r2_gputest_v3b.log $OUT/r2_bench_v2b.err $OUT/r2_bench_v3b.err $OUT/r2_cfg5_n1_p1000_cpu.err; do echo "== $f"; tail -2 $f; done
