#!/bin/bash
# First GPU contact: kernel numerics, GEMM throughput vs cuBLAS, SASS/ncu evidence.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_gemm test_gpu_elementwise test_gpu_attention; do
  timeout 600 python -m pytest tests/$f.py -x -q -m gpu > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
timeout 600 python tools/bench_gemm.py --tokens 32768 --out gpurun_out/gemm_bench.json > gpurun_out/gemm_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/summary.txt
tail -30 gpurun_out/gemm_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/gemm_prof \
  python tools/bench_gemm.py --quick --tokens 16384 --out gpurun_out/gemm_quick.json > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
