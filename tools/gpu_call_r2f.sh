#!/bin/bash
# round-2 GPU call F: the phase-scheduled serial stage (k_rate_ph): parity first, then A/B against the nested kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q > $O/r2f_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.txt
Q="timeout 300 python tools/quick_bench.py"
HMP3_RATE_MODE=nested $Q 4736 12 > $O/r2f_q_nested.txt 2>&1
$Q 4736 12 > $O/r2f_q_ph16.txt 2>&1
HMP3_RATE_PH_WARPS=8 $Q 4736 12 > $O/r2f_q_ph8.txt 2>&1
HMP3_RATE_PH_WARPS=12 $Q 4736 12 > $O/r2f_q_ph12.txt 2>&1
$Q 9472 12 > $O/r2f_q_ph16_9472.txt 2>&1
HMP3_RATE_MODE=nested $Q 9472 12 > $O/r2f_q_nested_9472.txt 2>&1
echo done
