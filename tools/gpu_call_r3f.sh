#!/bin/bash
# small batches: nested (one warp per stream) against phase-scheduled
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 100 python tools/quick_bench.py"
for n in 1250 2500; do
  $Q $n 30 2>&1 | grep -E "^run 1|rate_loop" | tail -2 > $O/r3f_ph_$n.txt
  HMP3_RATE_MODE=nested $Q $n 30 2>&1 | grep -E "^run 1|rate_loop" | tail -2 > $O/r3f_nested_$n.txt
done
HMP3_RATE_PH_SLOTS=18 $Q 1250 30 2>&1 | grep -E "^run 1|rate_loop" | tail -2 > $O/r3f_ph_s18_1250.txt
HMP3_RATE_PH_SLOTS=36 $Q 2500 30 2>&1 | grep -E "^run 1|rate_loop" | tail -2 > $O/r3f_ph_s36_2500.txt
echo done
