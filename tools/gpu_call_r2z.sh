#!/bin/bash
# round-2 GPU call Z: full captures of the Phase A / packing kernels of the final build + counters of k_rate_ph
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pack|k_polyphase|k_hybrid|k_psy_stage1|k_prepare$|k_psy_stage2|k_attack' -s 14 -c 7 -o $O/r2z2_others python tools/quick_bench.py 4736 12 > $O/r2z2.log 2>&1
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 3 -c 1 --csv --log-file $O/r2z2_ph_9472.csv python tools/quick_bench.py 9472 30 > $O/r2z2_a.log 2>&1
echo done
