#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 60 python tools/quick_bench.py"
HMP3_RATE_PH_OPTS=4 $Q 4736 4 > $O/r2k_dbg.txt 2>&1
HMP3_RATE_PH_OPTS=5 $Q 4736 4 > $O/r2k_nochain.txt 2>&1
HMP3_RATE_PH_OPTS=6 $Q 4736 4 > $O/r2k_noprefetch.txt 2>&1
HMP3_RATE_PH_OPTS=4 HMP3_RATE_PH_WARPS=1 $Q 4736 4 > $O/r2k_w1.txt 2>&1
echo done
