#!/bin/bash
# split-carrier kernels with the filter loads one chunk ahead: parity, then the 100-permutation step
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_join_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > $OUT/r2_gputest_pipe.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_pipe.log
timeout 300 python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p100_pipe.json 2> $OUT/r2_p100_pipe.err
GCRE_SC_PTS=0 timeout 300 python bench.py --n-perms 100 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p100_pipe_nopts.json 2> $OUT/r2_p100_pipe_nopts.err
tail -n 4 $OUT/r2_gputest_pipe.log
for f in p100_pipe p100_pipe_nopts; do python - $OUT/r2_$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    pl=d['per_level']
    print(sys.argv[1], "ms/step %.2f"%d['ms_per_step'], " ".join("%s:%s=%.2f"%(m[-1],k,v['kernel_ms']) for m in pl for k,v in pl[m].items()))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
