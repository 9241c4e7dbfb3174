#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests25.log 2>&1
echo "full gpu suite exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/tests25.log
python tools/exp_small.py 2>&1 | head -2
python bench.py --model vitl --steps 10 --warmup 3 > gpurun_out/bench_vitl3.log 2>&1; tail -1 gpurun_out/bench_vitl3.log | cut -c1-260
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
cat gpurun_out/summary.txt
