#!/bin/bash
# round-2 GPU call Q: coarser phase sets through chaining (opts 32 / 64), a lighter ncu capture with source counters
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 200 python tools/quick_bench.py"
for N in 9472 4736; do
for o in 0 32 64 96; do
  HMP3_RATE_PH_OPTS=$o $Q $N 12 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2q_opts${o}_$N.txt
done
done
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section Occupancy --section LaunchStats --clock-control none --import-source on -k regex:k_rate_ph -s 2 -c 1 -o $O/r2q_rate_ph python tools/quick_bench.py 4736 12 > $O/r2q_b.log 2>&1
echo done
