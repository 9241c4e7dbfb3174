#!/bin/bash
# final 8-GPU runs: ours (10B, ViT-L) then the reference arm (10B)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_10b_n8_final.log 2>&1
echo "ours 10b n8 exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_10b_n8_final.log | cut -c1-400
P=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --model vitl --steps 10 --warmup 3 > gpurun_out/bench_vitl_n8_final.log 2>&1
echo "ours vitl n8 exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_vitl_n8_final.log | cut -c1-300
P=$((20000 + RANDOM % 20000))
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_10b_n8_ref.log 2>&1
echo "ref 10b n8 exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_10b_n8_ref.log | cut -c1-400
cat gpurun_out/summary.txt
