#!/bin/bash
# 8 GPUs: validate the full stack at W=8 and get the headline number
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
P=$((20000 + RANDOM % 20000))
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_10b_n8.log 2>&1
echo "bench 10b n8 exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_10b_n8.log | cut -c1-1500
P=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --model vitl --steps 5 --warmup 3 > gpurun_out/bench_vitl_n8.log 2>&1
echo "bench vitl n8 exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_vitl_n8.log | cut -c1-900
cat gpurun_out/summary.txt
