#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 500 python -m pytest tests/test_gpu_boundary.py -q -k "rate_conversion" > $O/r3h_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r3h_pytest.txt
echo done
