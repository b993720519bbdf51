#!/bin/bash
# round-2 GPU call Y: full GPU suite + the bench line of the current build
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2y_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r2y_pytest.txt
[ $rc = 0 ] || { echo "gpu tests failed"; tail -30 $O/r2y_pytest.txt; exit 1; }
timeout 700 python bench.py > $O/r2y_bench.json 2> $O/r2y_bench.err
timeout 400 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2y_smoke.txt 2>&1
echo done
