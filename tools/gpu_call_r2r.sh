#!/bin/bash
# round-2 GPU call R: fewer serial-stage warps with the 80-register instance, so that Phase A / packing blocks co-reside
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 200 python tools/quick_bench.py"
for w in 24 22 20 18 16; do
  HMP3_RATE_PH_REGS80=1 HMP3_RATE_PH_WARPS=$w $Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2r_w${w}_9472.txt
done
HMP3_RATE_PH_REGS80=1 HMP3_RATE_PH_WARPS=20 HMP3_RATE_CARVEOUT=30 $Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2r_w20c30_9472.txt
# by-function profile: the 80-register instance with 20 warps leaves the profiler's patched code room for its registers
timeout 900 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section Occupancy --section LaunchStats --clock-control none --import-source on -k regex:k_rate_ph -s 2 -c 1 -o $O/r2r_rate_ph env HMP3_RATE_PH_REGS80=1 HMP3_RATE_PH_WARPS=20 python tools/quick_bench.py 9472 12 > $O/r2r_b.log 2>&1
echo done
