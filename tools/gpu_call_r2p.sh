#!/bin/bash
# round-2 GPU call P: hot/cold split of the serial state + L2 persistence window; full ncu capture of k_rate_ph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_encode.py -x -q > $O/r2p_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r2p_pytest.txt
[ $rc = 0 ] || { echo "gpu tests failed"; tail -30 $O/r2p_pytest.txt; exit 1; }
Q="timeout 200 python tools/quick_bench.py"
$Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2p_persist_9472.txt
HMP3_RATE_L2_PERSIST_MB=0 $Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2p_nopersist_9472.txt
HMP3_RATE_L2_PERSIST_MB=40 $Q 9472 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2p_persist40_9472.txt
$Q 4736 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2p_persist_4736.txt
HMP3_RATE_L2_PERSIST_MB=0 $Q 4736 30 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2p_nopersist_4736.txt
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 3 -c 1 --csv --log-file $O/r2p_ph_9472.csv python tools/quick_bench.py 9472 30 > $O/r2p_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rate_ph -s 3 -c 1 -o $O/r2p_rate_ph python tools/quick_bench.py 9472 30 > $O/r2p_b.log 2>&1
echo done
