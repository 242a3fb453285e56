#!/bin/bash
# split-carrier kernels (<= 512 permutations): parity, then the 100-permutation step
set -u
OUT=gpurun_out
python -m pytest tests -m gpu -q -x > $OUT/r2_gputest_sc3.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_sc3.log
GCRE_TRACE=1 python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p100_sc3.json 2> $OUT/r2_p100_sc3.err
for V in mb10; do
  GCRE_B200_LIB=$PWD/build_variants/lib_$V.so python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p100_sc3_$V.json 2> $OUT/r2_p100_sc3_$V.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p1000_sc3.json 2> $OUT/r2_p1000_sc3.err
tail -n 4 $OUT/r2_gputest_sc3.log
for f in p100_sc3 p100_sc3_mb10 p1000_sc3; do python - $OUT/r2_$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    pl=d['per_level']
    print(sys.argv[1], "ms/step %.2f"%d['ms_per_step'], " ".join("%s:%s=%.2f"%(m[-1],k,v['kernel_ms']) for m in pl for k,v in pl[m].items() if k in('2','3','4')))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
