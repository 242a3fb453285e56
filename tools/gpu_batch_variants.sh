#!/bin/bash
# Kernel variants built from working copies (build_variants/lib_*.so), one B200: parity subset, then bench with resident inputs.
set -u
OUT=gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_head.json 2> $OUT/r2_bench_head.err
for V in v6; do
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python -m pytest tests/test_join_gpu.py tests/test_fullsize_gpu.py -q -x -k "schedule_matches_oracle or golden or many_permutations or fullsize or agree or compose or first_rows" > $OUT/r2_gputest_$V.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_$V.log
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_$V.json 2> $OUT/r2_bench_$V.err
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_${V}_p100.json 2> $OUT/r2_bench_${V}_p100.err
done
for V in v6; do echo "== $V"; tail -n 2 $OUT/r2_gputest_$V.log 2>/dev/null; done
