#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_attention.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/tests20.log 2>&1
echo "tests exit $?" >> gpurun_out/summary.txt
tail -6 gpurun_out/tests20.log
timeout 600 python bench.py --model vitl --steps 10 --warmup 3 > gpurun_out/bench_vitl2.log 2>&1
tail -1 gpurun_out/bench_vitl2.log | cut -c1-330
timeout 900 python bench.py --model vit10b --steps 3 --warmup 3 > gpurun_out/bench_10b_3.log 2>&1
tail -1 gpurun_out/bench_10b_3.log | cut -c1-330
cat gpurun_out/summary.txt
