#!/bin/bash
# round-2 GPU call U: chunk length at 9472 streams per GPU
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
Q="timeout 200 python tools/quick_bench.py"
for g in 64 96 128 192 256; do
  HMP3_CHUNK_GRANULES=$g $Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror|failed" > $O/r2u_ng${g}_9472.txt
done
HMP3_RATE_L2_PERSIST_MB=72 $Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror|failed" > $O/r2u_persist72_9472.txt
echo done
