#!/bin/bash
# Final single-GPU artefacts again after the last kernel change (true scores by all lanes in the 1,024-permutation kernels):
# bench line, launch list, instruction counts, ncu --set full of the join launches.  Every ncu command repeats a command that
# exited 0 without ncu in this batch.
set -u
OUT=gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-strong"
$B > $OUT/r2_prof_plain.json 2> $OUT/r2_prof_plain.err
ncu --metrics $M --clock-control none -k regex:join_ -c 60 --csv --log-file $OUT/r2_counts_cfg3.csv $B > $OUT/r2_ncu_counts_cfg3.log 2>&1
python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_n1.json 2> $OUT/r2_bench_n1.err; echo "rc=$?" >> $OUT/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_step_launches.csv $B > $OUT/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:join_sparse_kernel -c 10 -o $OUT/r2_final_full \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_ncu_full.log 2>&1
tail -n 2 $OUT/r2_bench_n1.err
