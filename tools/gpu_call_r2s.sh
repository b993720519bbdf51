#!/bin/bash
# round-2 GPU call S: per-GPU shapes of C5's strong scaling (10000 clips over 2/4/8 GPUs = 5000/2500/1250 per GPU; no
# collective, so one GPU with that many clips is the per-GPU figure), source-level capture of k_rate_ph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for n in 5000 2500 1250; do
  timeout 400 python bench.py --clips-per-gpu $n --no-cpu-baseline --parity-streams 2 --steps 3 --warmup 3 > $O/r2s_bench_$n.json 2> $O/r2s_bench_$n.err
done
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section Occupancy --section LaunchStats --clock-control none --import-source on -k regex:k_rate_ph -s 2 -c 1 -o $O/r2s_rate_ph env HMP3_RATE_PH_REGS80=1 HMP3_RATE_PH_WARPS=20 python tools/quick_bench.py 9472 12 > $O/r2s_b.log 2>&1
echo done
