#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for f in test_gpu_attention test_gpu_engine; do
  timeout 900 python -m pytest tests/$f.py -x -q -m gpu > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
timeout 600 python tools/bench_gemm.py --tokens 32768 --out gpurun_out/gemm_bench2.json > gpurun_out/gemm_bench2.log 2>&1
echo "gemm bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --model vitl --steps 5 --warmup 3 > gpurun_out/bench_vitl.log 2>&1
echo "bench vitl exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_vitl.log
timeout 900 python bench.py --model vit10b --num_blocks 4 --steps 3 --warmup 3 > gpurun_out/bench_10b_4blk.log 2>&1
echo "bench 10b-4blk exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_10b_4blk.log
timeout 1200 python bench.py --model vit10b --steps 3 --warmup 3 > gpurun_out/bench_10b.log 2>&1
echo "bench 10b exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/bench_10b.log
cat gpurun_out/summary.txt
