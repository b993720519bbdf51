#!/bin/bash
# round-2 GPU call W: CLI -A cases against the reference CLI; launch list of one pipeline run (9472 x 30 s)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_wav_formats.py -x -q -k "up_converted or fuzz or option_strings" > $O/r2w_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r2w_pytest.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2w_launches.csv python tools/quick_bench.py 9472 30 > $O/r2w_c.log 2>&1
echo done
