#!/bin/bash
# exposed-communication probe at N=2 (ViT-10B)
mkdir -p gpurun_out
P=$((20000 + RANDOM % 20000))
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 2 --warmup 3 --no_e2e --no_full_ckpt_probe 2>&1 | tail -1 | tee gpurun_out/exposed_n2.log
