#!/bin/bash
# round-2 GPU call B: full GPU test suite on the int16 / segmented-max / line-parallel refit build, quick bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err
echo done
