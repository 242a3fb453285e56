#!/bin/bash
# Round-2 measurement batch for one B200 (run through gpurun from the repo root).
set -u
OUT=gpurun_out
python -m pytest tests -m gpu -q > $OUT/r2_gputest2.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest2.log
python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_b.json 2> $OUT/r2_bench_b.err; echo "bench rc=$?" >> $OUT/r2_bench_b.err
# occupancy variants of the sparse kernels (resident CTAs per SM for methods 1 / 2), same bench, resident inputs only
for V in 8_mb2_6 8_mb2_7 10_mb2_8 12_mb2_8 10_mb2_7; do
  GCRE_B200_LIB=$PWD/build_variants/lib_mb1_$V.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_var_$V.json 2> $OUT/r2_bench_var_$V.err
done
bash tools/gpu_batch_profile.sh > $OUT/r2_profile_batch.log 2>&1
# per-kernel split of the opt-in screening path (launch list only)
GCRE_SCREEN=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_step_launches_screen.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r2_ncu_launches_screen.log 2>&1
# v3 kernels (range-bound pre-check in place of the per-permutation compare; built from the working copy): parity first, then timing
GCRE_B200_LIB=$PWD/build_variants/lib_v3_8_8.so python -m pytest tests/test_join_gpu.py -q -x -k "schedule_matches_oracle or golden or many_permutations or dense_carrier" > $OUT/r2_gputest_v3.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_v3.log
GCRE_B200_LIB=$PWD/build_variants/lib_v3_8_8.so python bench.py --steps 10 --warmup 3 --no-e2e > $OUT/r2_bench_v3_8_8.json 2> $OUT/r2_bench_v3_8_8.err
for V in 10_8 12_8 8_6; do
  GCRE_B200_LIB=$PWD/build_variants/lib_v3_$V.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_v3_$V.json 2> $OUT/r2_bench_v3_$V.err
done
# BASELINE config 5: permutation sweep at 50,000 patients, path length 4, one GPU; the CPU leg (sub-shape with identical W) beside the 1,000-permutation point
for P in 100 1000 10000; do
  EXTRA="--no-cpu-baseline"; [ "$P" = "1000" ] && EXTRA="--cpu-seconds 8"
  python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms $P --steps 3 --warmup 3 --no-e2e $EXTRA > $OUT/r2_cfg5_n1_p$P.json 2> $OUT/r2_cfg5_n1_p$P.err
  echo "cfg5 p$P rc=$?" >> $OUT/r2_cfg5_n1_p$P.err
done
python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 100000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2_cfg5_n1_p100000.json 2> $OUT/r2_cfg5_n1_p100000.err
echo "cfg5 p100000 rc=$?" >> $OUT/r2_cfg5_n1_p100000.err
tail -3 $OUT/r2_gputest2.log $OUT/r2_bench_b.err $OUT/r2_cfg5_n1_p*.err
