#!/bin/bash
# 2 GPUs: fused RS+AdamW test, reference arm at N=2 for the head-to-head
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "adamw_fused" > gpurun_out/test_gpu_multi2.log 2>&1
echo "fused adamw test exit $?" >> gpurun_out/summary.txt
tail -8 gpurun_out/test_gpu_multi2.log
P=$((20000 + RANDOM % 20000))
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_10b_n2_ref.log 2>&1
echo "bench 10b n2 ref exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_10b_n2_ref.log | cut -c1-1500
cat gpurun_out/summary.txt
