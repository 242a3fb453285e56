#!/bin/bash
# quick check of the adaptive end-to-end warm-up in bench.py
set -u
OUT=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong --e2e-steps 10 > $OUT/r2_e2e_adaptive.json 2> $OUT/r2_e2e_adaptive.err; echo "rc=$?" >> $OUT/r2_e2e_adaptive.err
tail -n 2 $OUT/r2_e2e_adaptive.err
