#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/raster.log
run() {  # group hintA hintB shape
  export B200_GEMM_GROUP_N=$1 B200_GEMM_HINT_A=$2 B200_GEMM_HINT_B=$3 EXP_SHAPE=$4
  timeout 120 python tools/exp_raster.py >> gpurun_out/raster.log 2>&1
  EXP_ITERS=1 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
     -k regex:gemm_bf16 -s 3 -c 1 --csv python tools/exp_raster.py 2>/dev/null | grep -E "dram__bytes|hit_rate|duration" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ';' >> gpurun_out/raster.log
  echo >> gpurun_out/raster.log
}
for g in 4 8 16 30 60; do run $g 0 0 qkv; done
run 8 1 2 qkv
run 16 1 2 qkv
run 8 2 2 qkv
run 8 0 0 fc2
run 4 0 0 fc2
run 20 0 0 fc2
run 8 1 2 fc2
cat gpurun_out/raster.log
