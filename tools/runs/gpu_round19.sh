#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_vitl.csv \
  python bench.py --model vitl --steps 1 --warmup 3 --no_e2e --cuda_graph 0 > gpurun_out/launches_vitl.log 2>&1
echo "exit $?"; wc -l gpurun_out/launches_vitl.csv
