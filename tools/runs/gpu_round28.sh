#!/bin/bash
# memory-aware activation keeping: GPU tests, then ViT-10B at N=1 (tight memory) 
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 4 --warmup 3 2>&1 | tail -3 | tee gpurun_out/keep_n1.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
