#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/tests18.log 2>&1
echo "engine tests exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/tests18.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/tests18m.log 2>&1
echo "multi tests exit $?" >> gpurun_out/summary.txt
tail -15 gpurun_out/tests18m.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --model vitl --steps 10 --warmup 3 > gpurun_out/bench_vitl_graph.log 2>&1
echo "vitl graph exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_vitl_graph.log | cut -c1-500
P=$((20000 + RANDOM % 20000))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --model vitl --steps 10 --warmup 3 > gpurun_out/bench_vitl_graph_n2.log 2>&1
echo "vitl graph n2 exit $?" >> gpurun_out/summary.txt
tail -1 gpurun_out/bench_vitl_graph_n2.log | cut -c1-500
cat gpurun_out/summary.txt
