#!/bin/bash
# round-2 GPU call H: the phase-scheduled serial stage, build variants (optimisation level, warps per block)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 300 python tools/quick_bench.py"
for v in main ph_o3o3 ph_o1o3 ph_w24 ph_w32 ph_w24o3 ph_w32o3; do
  L=$PWD/hmp3_b200/_lib/var_$v.so; [ $v = main ] && L=$PWD/hmp3_b200/_lib/libhmp3_b200.so
  for n in 4736 9472; do
    HMP3_B200_LIB=$L $Q $n 12 2>&1 | grep -E "^run|rate_loop" > $O/r2h_${v}_$n.txt
  done
done
echo done
