#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/raster2.log
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -x -q -m gpu > gpurun_out/test_gpu_gemm_clc.log 2>&1
echo "gemm tests (CLC) exit $?" >> gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_gemm_clc.log
run() {  # clc group shape
  export B200_GEMM_CLC=$1 B200_GEMM_GROUP_N=$2 EXP_SHAPE=$3
  timeout 120 python tools/exp_raster.py >> gpurun_out/raster2.log 2>&1
  EXP_ITERS=1 timeout 200 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
     -k regex:gemm_bf16 -s 3 -c 1 --csv python tools/exp_raster.py 2>/dev/null | grep -E "dram__bytes|hit_rate|duration" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ';' >> gpurun_out/raster2.log
  echo " clc=$1" >> gpurun_out/raster2.log
}
run 1 8 qkv; run 1 16 qkv; run 0 8 qkv; run 1 8 fc2; run 1 4 fc2; run 1 16 fc2; run 0 8 fc2
cat gpurun_out/raster2.log
unset B200_GEMM_CLC B200_GEMM_GROUP_N EXP_SHAPE
timeout 300 python tools/exp_sustained.py > gpurun_out/sustained2.log 2>&1
cat gpurun_out/sustained2.log
cat gpurun_out/summary.txt
