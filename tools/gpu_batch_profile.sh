#!/bin/bash
# Profiling batch for one B200 (run through gpurun from the repo root): ncu launch list + full captures of one step, sanitizer runs.
# Every ncu command repeats a command that has already exited 0 without ncu in this batch; numbers printed under ncu are not bench values.
set -u
OUT=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$B > $OUT/r2_prof_plain.json 2> $OUT/r2_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_step_launches.csv $B > $OUT/r2_ncu_launches.log 2>&1
# one step without warm-up: 2 methods x (levels 1a,1b,2,3 + seed, screen, retry of level 4) = 14 join launches
ncu --set full --clock-control none --import-source on -k regex:"join_screen_kernel|join_sparse_kernel" -c 14 -o $OUT/r2_step_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/r2_ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_join_gpu.py -q -x \
    -k "corner or topk_launch or (schedule_matches_oracle and sparse_screen) or (dense_carrier_rows and 96)" > $OUT/r2_memcheck.log 2>&1
echo "memcheck rc=$?" >> $OUT/r2_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_join_gpu.py -q -x \
    -k "corner or topk_launch or (schedule_matches_oracle and sparse_screen and shape3)" > $OUT/r2_racecheck.log 2>&1
echo "racecheck rc=$?" >> $OUT/r2_racecheck.log
tail -3 $OUT/r2_memcheck.log $OUT/r2_racecheck.log
