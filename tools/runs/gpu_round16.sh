#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/tests16.log 2>&1
echo "tests exit $?" >> gpurun_out/summary.txt
tail -4 gpurun_out/tests16.log
for c in dgrad_fc2_dgelu dgrad_fc2 wgrad_fc2; do EXP_CASE=$c timeout 120 python tools/exp_wgrad.py; done > gpurun_out/wgrad2.log 2>&1
cat gpurun_out/wgrad2.log
timeout 300 python tools/prof_kernels.py; timeout 300 ncu --metrics gpu__time_duration.sum -k regex:"gemm_bf16" -s 6 -c 3 --csv python tools/prof_kernels.py 2>/dev/null | grep duration | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-160 > gpurun_out/attn_gemm_times.log
cat gpurun_out/attn_gemm_times.log
timeout 900 python bench.py --model vit10b --steps 3 --warmup 3 > gpurun_out/bench_10b_e8.log 2>&1
tail -1 gpurun_out/bench_10b_e8.log | cut -c1-420
cat gpurun_out/summary.txt
