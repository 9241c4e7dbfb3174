#!/bin/bash
# memory-aware activation keeping at N=2 (memory-limited K) : GPU tests + ViT-10B + ViT-L
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 4 --warmup 3 2>&1 | tail -1 | tee gpurun_out/keep_n2.log
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 3 --model vitl 2>&1 | tail -1 | tee gpurun_out/keep_n2_vitl.log
