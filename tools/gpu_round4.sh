#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_elementwise.py -x -q -m gpu > gpurun_out/test_gpu_engine.log 2>&1
echo "tests exit $?" >> gpurun_out/summary.txt
tail -3 gpurun_out/test_gpu_engine.log
# per-kernel device time of one ViT-10B (2 blocks) training step: shares, not absolutes
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_10b_2blk.csv \
  python bench.py --model vit10b --num_blocks 2 --steps 1 --warmup 3 --no_e2e > gpurun_out/launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
