#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_ln.log; : > $L
timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -12 >> $L
B200_LN_STREAM=0 timeout 120 python tools/exp_ln_trace.py 2>&1 | tail -2 >> $L
timeout 200 python tools/exp_ln_trace.py --trace 2>&1 | tail -14 >> $L
cat $L
