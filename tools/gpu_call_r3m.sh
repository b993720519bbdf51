#!/bin/bash
# round-2 final GPU call: the whole GPU suite, then smoke()
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 275 python -m pytest tests -m gpu -x -q --durations=8 > $O/r3m_pytest.txt 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/r3m_pytest.txt
tail -14 $O/r3m_pytest.txt
timeout 45 python -c "import __graft_entry__ as g; g.smoke()" > $O/r3m_smoke.txt 2>&1; echo "smoke rc=$?" >> $O/r3m_smoke.txt
tail -2 $O/r3m_smoke.txt
