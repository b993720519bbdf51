#!/bin/bash
# round-2 GPU call O: counters + full capture of the phase-scheduled serial stage at the bench shape (64 streams per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 3 -c 1 --csv --log-file $O/r2o_ph_9472.csv python tools/quick_bench.py 9472 30 > $O/r2o_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rate_ph -s 3 -c 1 -o $O/r2o_rate_ph python tools/quick_bench.py 9472 30 > $O/r2o_b.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2o_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --parity-streams 1 > $O/r2o_c.log 2>&1
echo done
