#!/bin/bash
# round-2 GPU call J: guarded re-run of call I (every step under a short timeout; stop at the first failure)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 240 python -m pytest tests/test_gpu_encode.py -x -q > $O/r2j_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r2j_pytest.txt
[ $rc = 0 ] || { echo "parity failed"; exit 1; }
Q="timeout 100 python tools/quick_bench.py"
$Q 4736 12 > $O/r2j_main_4736.txt 2>&1 || { echo "bench failed"; tail -5 $O/r2j_main_4736.txt; exit 1; }
for v in main ph_w24 ph_w32; do
  L=$PWD/hmp3_b200/_lib/var_$v.so; [ $v = main ] && L=$PWD/hmp3_b200/_lib/libhmp3_b200.so
  for n in 4736 9472; do
    HMP3_B200_LIB=$L $Q $n 12 2>&1 | grep -E "^run|rate_loop|rror" > $O/r2j_${v}_$n.txt
  done
done
HMP3_RATE_PH_WARPS=12 $Q 4736 12 2>&1 | grep -E "^run|rate_loop" > $O/r2j_main_w12_4736.txt
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 2 -c 1 --csv --log-file $O/r2j_ph16_4736.csv python tools/quick_bench.py 4736 12 > $O/r2j_a.log 2>&1
echo done
