#!/bin/bash
# round-2 final GPU call: full GPU suite, smoke, bench line, counters of k_rate_ph, launch list, ncu of the other kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r3c_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r3c_pytest.txt
[ $rc = 0 ] || { echo "gpu tests failed"; tail -30 $O/r3c_pytest.txt; exit 1; }
timeout 400 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r3c_smoke.txt 2>&1
timeout 700 python bench.py > $O/r3c_bench.json 2> $O/r3c_bench.err
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_issue_stalled_short_scoreboard.sum,smsp__warps_issue_stalled_wait.sum,smsp__warps_issue_stalled_sleeping.sum,smsp__warps_issue_stalled_barrier.sum,smsp__warps_issue_stalled_membar.sum,smsp__warps_issue_stalled_math_pipe_throttle.sum,smsp__warps_issue_stalled_lg_throttle.sum,smsp__warps_issue_stalled_dispatch_stall.sum,smsp__warps_issue_stalled_not_selected.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_rate_ph -s 3 -c 1 --csv --log-file $O/r3c_ph_9472.csv python tools/quick_bench.py 9472 30 > $O/r3c_a.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r3c_launches.csv python tools/quick_bench.py 9472 30 > $O/r3c_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pack|k_polyphase|k_hybrid|k_psy_stage1|k_prepare$|k_psy_stage2|k_attack' -s 14 -c 7 -o $O/r3c_others python tools/quick_bench.py 4736 12 > $O/r3c_z.log 2>&1
echo done
