#!/bin/bash
# 1,024-permutation kernels with the true scores worked off by all lanes every 32 pairs: parity subset, then the default step
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_join_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > $OUT/r2_gputest_defer.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_gputest_defer.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > $OUT/r2_p1000_defer.json 2> $OUT/r2_p1000_defer.err
tail -n 3 $OUT/r2_gputest_defer.log
python - $OUT/r2_p1000_defer.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
pl=d['per_level']
print("ms/step %.2f"%d['ms_per_step'], " ".join("%s:%s=%.2f"%(m[-1],k,v['kernel_ms']) for m in pl for k,v in pl[m].items()))
PY
