#!/bin/bash
# N=8: memory-aware activation keeping (all 32 blocks fit), ViT-10B + ViT-L
mkdir -p gpurun_out
P=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 4 --warmup 3 2>&1 | tail -1 | tee gpurun_out/keep_n8.log
P=$((20000 + RANDOM % 20000))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 3 --model vitl 2>&1 | tail -1 | tee gpurun_out/keep_n8_vitl.log
