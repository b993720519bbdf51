#!/bin/bash
# round-2 GPU call C: convoy experiments (block-wide barriers per frame / granule, 4..32 warps per block) + the main build
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "encode or boundary" > $O/r2c_pytest.txt 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.txt
for v in main b32 b32f b32g b16g b8g b4g; do
  L=$PWD/hmp3_b200/_lib/var_$v.so; [ $v = main ] && L=$PWD/hmp3_b200/_lib/libhmp3_b200.so
  HMP3_B200_LIB=$L timeout 400 python bench.py --no-cpu-baseline --parity-streams 4 --steps 2 --warmup 2 > $O/r2c_$v.json 2> $O/r2c_$v.err
done
echo done
