#!/bin/bash
# Round-2 final measurement batch for one B200 (run through gpurun from the repo root).  Every ncu command repeats a command that
# has already exited 0 without ncu in this batch; numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"
python -m pytest tests -m gpu -q > $OUT/r2_gputest_final.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_final.log
python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_n1.json 2> $OUT/r2_bench_n1.err; echo "rc=$?" >> $OUT/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/r2_bench_reference_arm.json 2> $OUT/r2_bench_reference_arm.err; echo "rc=$?" >> $OUT/r2_bench_reference_arm.err
GCRE_BENCH_DEBUG=1 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 6 > $OUT/r2_e2e_debug.json 2> $OUT/r2_e2e_debug.err
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$B > $OUT/r2_prof_plain.json 2> $OUT/r2_prof_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_step_launches.csv $B > $OUT/r2_ncu_launches.log 2>&1
ncu --metrics $M --clock-control none -k regex:join_ -c 60 --csv --log-file $OUT/r2_counts_cfg3.csv $B > $OUT/r2_ncu_counts_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:join_sparse_kernel -c 10 -o $OUT/r2_final_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/r2_ncu_full.log 2>&1
# other BASELINE shapes with this build
python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_config3_p100.json 2> $OUT/r2_bench_config3_p100.err
python bench.py --path-length 5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/r2_bench_config3_len5.json 2> $OUT/r2_bench_config3_len5.err
for P in 100 1000 10000; do
  EXTRA="--no-cpu-baseline"; [ "$P" = "1000" ] && EXTRA="--cpu-seconds 8"
  python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms $P --steps 3 --warmup 3 --no-e2e $EXTRA > $OUT/r2_config5_n1_p$P.json 2> $OUT/r2_config5_n1_p$P.err
  echo "rc=$?" >> $OUT/r2_config5_n1_p$P.err
done
python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 100000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/r2_config5_n1_p100000.json 2> $OUT/r2_config5_n1_p100000.err
ncu --metrics $M --clock-control none -k regex:join_ -c 60 --csv --log-file $OUT/r2_counts_cfg3_p100.csv \
    python bench.py --n-perms 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r2_ncu_counts_cfg3_p100.log 2>&1
ncu --metrics $M --clock-control none -k regex:join_ -c 60 --csv --log-file $OUT/r2_counts_cfg5_p100.csv \
    python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 100 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r2_ncu_counts_cfg5_p100.log 2>&1
ncu --metrics $M --clock-control none -k regex:join_ -c 60 --csv --log-file $OUT/r2_counts_cfg5_p1000.csv \
    python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 1000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/r2_ncu_counts_cfg5_p1000.log 2>&1
for f in $OUT/r2_gputest_final.log $OUT/r2_bench_n1.err $OUT/r2_bench_reference_arm.err $OUT/r2_config5_n1_p1000.err; do echo "== $f"; tail -n 2 $f; done
