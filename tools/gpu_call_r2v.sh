#!/bin/bash
# round-2 GPU call V (2 GPUs): the weak-scaling bench line at N = 2 with the phase-scheduled serial stage
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4 > $O/r2v_gpus.txt; free -g | head -2 >> $O/r2v_gpus.txt; nproc >> $O/r2v_gpus.txt
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2v_bench_n2.json 2> $O/r2v_bench_n2.err
echo done
