#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
HMP3_RATE_PH_OPTS=4 HMP3_RATE_PH_SLOTS=8 timeout 60 python tools/quick_bench.py 64 2 > $O/r2l_dbg64.txt 2>&1
echo done
