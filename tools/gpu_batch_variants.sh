#!/bin/bash
# What the driver runs at round end, on the committed tree: GPU tests, smoke(), the default bench line and the reference arm's start
set -u
OUT=gpurun_out
python -m pytest tests -x -q -m gpu > $OUT/r2_head_gputest.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_head_gputest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r2_head_smoke.log 2>&1; echo "rc=$?" >> $OUT/r2_head_smoke.log
( time python bench.py --gpus 1 --steps 5 --warmup 3 ) > $OUT/r2_head_bench.json 2> $OUT/r2_head_bench.err; echo "rc=$?" >> $OUT/r2_head_bench.err
tail -n 3 $OUT/r2_head_gputest.log; tail -n 3 $OUT/r2_head_smoke.log; tail -n 6 $OUT/r2_head_bench.err | cut -c1-200; grep -c nvcc $OUT/r2_head_bench.err
