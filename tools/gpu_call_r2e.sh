#!/bin/bash
# round-2 GPU call E: the tensor-core polyphase (first run: guarded by short timeouts), per-function profile of k_rate
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_polymm.py -q -x > $O/r2e_polymm_test.txt 2>&1; echo "rc=$?" >> $O/r2e_polymm_test.txt
if grep -q "passed" $O/r2e_polymm_test.txt; then
  timeout 900 python tools/polymm_eval.py 2048 10 > $O/r2e_polymm_eval.json 2> $O/r2e_polymm_eval.err
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rate$ -s 3 -c 1 -o $O/r2e_rate python tools/quick_bench.py 4736 30 > $O/r2e_rate_ncu.log 2>&1
echo done
