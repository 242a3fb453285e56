#!/bin/bash
# e2e robustness check (one B200): host packer with dynamic chunking, 16 vs 12 pack threads
set -u
OUT=gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 30 > $OUT/r2_e2e_dyn16.json 2> $OUT/r2_e2e_dyn16.err
GCRE_HOST_PACK_THREADS=12 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 30 > $OUT/r2_e2e_dyn12.json 2> $OUT/r2_e2e_dyn12.err
grep "e2e ms per step" $OUT/r2_e2e_dyn16.err $OUT/r2_e2e_dyn12.err
