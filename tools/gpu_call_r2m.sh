#!/bin/bash
# round-2 GPU call M: scheduling policy (most waiting / pipeline order), streams per SM
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 100 python tools/quick_bench.py"
run() { name=$1; shift; env "$@" $Q $N 12 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2m_${name}_$N.txt; }
for N in 4736 9472; do
  run popular HMP3_RATE_PH_OPTS=0
  run cyclic HMP3_RATE_PH_OPTS=8
  run cyclic_w12 HMP3_RATE_PH_OPTS=8 HMP3_RATE_PH_WARPS=12
  run cyclic_w8 HMP3_RATE_PH_OPTS=8 HMP3_RATE_PH_WARPS=8
done
N=9472; run popular_s128 HMP3_RATE_PH_OPTS=0 HMP3_RATE_PH_SLOTS=128
N=9472; run cyclic_s128 HMP3_RATE_PH_OPTS=8 HMP3_RATE_PH_SLOTS=128
N=4736; run popular_s64 HMP3_RATE_PH_OPTS=0 HMP3_RATE_PH_SLOTS=64
N=4736; run cyclic_s64 HMP3_RATE_PH_OPTS=8 HMP3_RATE_PH_SLOTS=64
echo done
