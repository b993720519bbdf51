#!/bin/bash
# round-2 GPU call N: full GPU test suite with the phase-scheduled serial stage, bench lines (9472 / 4736 / 10000 clips)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2n_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r2n_pytest.txt
[ $rc = 0 ] || { echo "gpu tests failed"; tail -30 $O/r2n_pytest.txt; exit 1; }
timeout 700 python bench.py > $O/r2n_bench.json 2> $O/r2n_bench.err
timeout 500 python bench.py --clips-per-gpu 4736 --no-cpu-baseline --parity-streams 2 > $O/r2n_bench_4736.json 2> $O/r2n_bench_4736.err
HMP3_RATE_MODE=nested timeout 500 python bench.py --no-cpu-baseline --parity-streams 2 > $O/r2n_bench_nested.json 2> $O/r2n_bench_nested.err
timeout 500 python bench.py --total-clips 10000 --no-cpu-baseline --parity-streams 2 > $O/r2n_strong1.json 2> $O/r2n_strong1.err
echo done
