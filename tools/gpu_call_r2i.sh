#!/bin/bash
# round-2 GPU call I: finer phases + bit-mask scheduler + frame-input prefetch; parity, then timing and counters
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_encode.py -x -q > $O/r2i_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest.txt
Q="timeout 300 python tools/quick_bench.py"
for v in main ph_w24 ph_w32 ph_o1o3; do
  L=$PWD/hmp3_b200/_lib/var_$v.so; [ $v = main ] && L=$PWD/hmp3_b200/_lib/libhmp3_b200.so
  for n in 4736 9472; do
    HMP3_B200_LIB=$L $Q $n 12 2>&1 | grep -E "^run|rate_loop" > $O/r2i_${v}_$n.txt
  done
done
HMP3_RATE_PH_WARPS=12 $Q 4736 12 2>&1 | grep -E "^run|rate_loop" > $O/r2i_main_w12_4736.txt
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 2 -c 1 --csv --log-file $O/r2i_ph16_4736.csv python tools/quick_bench.py 4736 12 > $O/r2i_a.log 2>&1
echo done
