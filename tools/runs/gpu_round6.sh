#!/bin/bash
# 2 GPUs: AG-fused GEMM + multi-GPU tests + ViT-10B scaling point
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s > gpurun_out/test_gpu_multi.log 2>&1
echo "multi tests exit $?" >> gpurun_out/summary.txt
tail -12 gpurun_out/test_gpu_multi.log
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_10b_n2.log 2>&1
echo "bench 10b n2 exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_10b_n2.log | cut -c1-1200
cat gpurun_out/summary.txt
