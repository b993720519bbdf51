#!/bin/bash
# round-2 GPU call T: L1 prefetch of the rows the streaming loops of the serial stage walk
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_encode.py -x -q > $O/r2t_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r2t_pytest.txt
[ $rc = 0 ] || { echo "gpu tests failed"; tail -30 $O/r2t_pytest.txt; exit 1; }
Q="timeout 200 python tools/quick_bench.py"
$Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2t_9472.txt
$Q 4736 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2t_4736.txt
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 3 -c 1 --csv --log-file $O/r2t_ph_9472.csv python tools/quick_bench.py 9472 30 > $O/r2t_a.log 2>&1
echo done
