#!/bin/bash
# round-2 GPU call A: tests, the new bench line, strong scaling at N=1, the W=16 build, instruction-cache counters
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/r2a_gpu.txt
free -g > $O/r2a_mem.txt; nproc >> $O/r2a_mem.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.txt
timeout 600 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err
timeout 600 python bench.py --total-clips 10000 --no-cpu-baseline --parity-streams 2 > $O/r2a_strong1.json 2> $O/r2a_strong1.err
HMP3_B200_LIB=$PWD/hmp3_b200/_lib/var_w16.so timeout 600 python bench.py --total-clips 10000 --no-cpu-baseline --parity-streams 2 > $O/r2a_strong1_w16.json 2> $O/r2a_strong1_w16.err
HMP3_B200_LIB=$PWD/hmp3_b200/_lib/var_w16.so timeout 600 python bench.py --no-cpu-baseline --parity-streams 2 > $O/r2a_w16.json 2> $O/r2a_w16.err
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_miss_tag_hit.sum,sm__icc_requests_lookup_miss_tag_miss.sum,sm__icc_requests_lookup_miss_tag_unavailable.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_hit.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,gcc__gcc2xbar_requests_type_instruction.sum,smsp__warps_issue_stalled_no_instruction.sum,smsp__warps_issue_stalled_branch_resolving.sum,smsp__warps_issue_stalled_long_scoreboard.sum,smsp__warps_active.sum,smsp__inst_executed.sum,smsp__issue_active.sum,sm__cycles_active.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_rate$ -s 2 -c 2 --csv --log-file $O/r2a_icache.csv python tools/quick_bench.py 4736 12 > $O/r2a_icache.log 2>&1
echo done
