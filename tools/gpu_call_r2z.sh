#!/bin/bash
# round-2 GPU call Z: full captures of the Phase A / packing kernels of the current build
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pack|k_polyphase|k_hybrid|k_psy_stage1|k_prepare$' -s 10 -c 5 -o $O/r2z_others python tools/quick_bench.py 4736 12 > $O/r2z.log 2>&1
echo done
