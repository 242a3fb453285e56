#!/bin/bash
# Round-end check of the committed tree on one B200: the whole GPU test suite and smoke()
set -u
OUT=gpurun_out
python -m pytest tests -x -q -m gpu > $OUT/r2_head_gputest.log 2>&1; echo "pytest rc=$?" >> $OUT/r2_head_gputest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r2_head_smoke.log 2>&1; echo "rc=$?" >> $OUT/r2_head_smoke.log
tail -n 3 $OUT/r2_head_gputest.log; tail -n 3 $OUT/r2_head_smoke.log
