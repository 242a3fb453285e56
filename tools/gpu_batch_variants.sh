#!/bin/bash
# method 2 at <= 128 permutations: both halves of a small pair in one pass - parity, then the 100-permutation step
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r2_gputest_both.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_both.log
timeout 300 python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p100_both.json 2> $OUT/r2_p100_both.err
timeout 300 python bench.py --n-cases 25000 --n-ctrls 25000 --n-perms 100 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-strong > $OUT/r2_cfg5_p100_both.json 2> $OUT/r2_cfg5_p100_both.err
tail -n 4 $OUT/r2_gputest_both.log
for f in p100_both cfg5_p100_both; do python - $OUT/r2_$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    pl=d['per_level']
    print(sys.argv[1], "ms/step %.2f"%d['ms_per_step'], " ".join("%s:%s=%.2f"%(m[-1],k,v['kernel_ms']) for m in pl for k,v in pl[m].items()))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
