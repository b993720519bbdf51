#!/bin/bash
# round-2 GPU call G: counters of the phase-scheduled serial stage
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_miss_tag_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 2 -c 1 --csv --log-file $O/r2g_ph16_4736.csv python tools/quick_bench.py 4736 12 > $O/r2g_a.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 2 -c 1 --csv --log-file $O/r2g_ph16_9472.csv python tools/quick_bench.py 9472 12 > $O/r2g_b.log 2>&1
HMP3_RATE_MODE=nested timeout 600 ncu --metrics $M --clock-control none -k regex:k_rate$ -s 2 -c 1 --csv --log-file $O/r2g_nested_4736.csv python tools/quick_bench.py 4736 12 > $O/r2g_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rate_ph -s 2 -c 1 -o $O/r2g_ph16_9472 python tools/quick_bench.py 9472 12 > $O/r2g_d.log 2>&1
echo done
