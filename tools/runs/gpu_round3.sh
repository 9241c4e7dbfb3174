#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/test_gpu_engine.log 2>&1
echo "test_gpu_engine exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_engine.log
timeout 600 python bench.py --model vitl --steps 5 --warmup 3 > gpurun_out/bench_vitl.log 2>&1
echo "bench vitl exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_vitl.log
timeout 600 python bench.py --impl reference --model vitl --steps 5 --warmup 3 > gpurun_out/bench_vitl_ref.log 2>&1
echo "bench vitl ref exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_vitl_ref.log
timeout 900 python bench.py --model vit10b --num_blocks 4 --steps 3 --warmup 3 > gpurun_out/bench_10b_4blk.log 2>&1
echo "bench 10b-4blk exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_10b_4blk.log
timeout 1200 python bench.py --model vit10b --steps 3 --warmup 3 > gpurun_out/bench_10b.log 2>&1
echo "bench 10b exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_10b.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/summary.txt
timeout 1500 python bench.py --impl reference --model vit10b --steps 3 --warmup 3 > gpurun_out/bench_10b_ref.log 2>&1
echo "bench 10b ref exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/bench_10b_ref.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -o gpurun_out/gemm_prof2 \
  python tools/bench_gemm.py --quick --tokens 32768 --out gpurun_out/gemm_quick.json > gpurun_out/ncu2.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
