#!/bin/bash
# round-2 GPU call D: fewer resident serial-stage warps so that Phase A / packing blocks can co-run on every SM
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
B="python bench.py --no-cpu-baseline --parity-streams 2 --steps 2 --warmup 2"
timeout 400 $B --clips-per-gpu 3552 > $O/r2d_3552.json 2> $O/r2d_3552.err
HMP3_PHASEA_CARVEOUT=40 timeout 400 $B --clips-per-gpu 3552 > $O/r2d_3552_c40.json 2> $O/r2d_3552_c40.err
HMP3_PHASEA_CARVEOUT=40 timeout 400 $B --clips-per-gpu 4144 > $O/r2d_4144_c40.json 2> $O/r2d_4144_c40.err
HMP3_PHASEA_CARVEOUT=40 timeout 400 $B --clips-per-gpu 4736 > $O/r2d_4736_c40.json 2> $O/r2d_4736_c40.err
HMP3_PHASEA_CARVEOUT=40 HMP3_CHUNK_SETS=3 timeout 400 $B --clips-per-gpu 3552 > $O/r2d_3552_c40_s3.json 2> $O/r2d_3552_c40_s3.err
HMP3_PHASEA_CARVEOUT=40 HMP3_NO_STREAM_PRIO=1 timeout 400 $B --clips-per-gpu 3552 > $O/r2d_3552_c40_np.json 2> $O/r2d_3552_c40_np.err
echo done
