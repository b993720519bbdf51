#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_encode.py -x -q -k "tap or scheduler" > $O/r3e_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r3e_pytest.txt
echo done
