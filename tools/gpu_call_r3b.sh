#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_analysis.py tests/test_gpu_encode.py -x -q > $O/r3b_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r3b_pytest.txt
[ $rc = 0 ] || { echo "gpu tests failed"; tail -30 $O/r3b_pytest.txt; exit 1; }
HMP3_SERIALIZE=1 timeout 200 python tools/quick_bench.py 9472 30 > $O/r3b_serial_9472.txt 2>&1
timeout 200 python tools/quick_bench.py 9472 30 > $O/r3b_9472.txt 2>&1
echo done
