#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_bwd_trace.log; : > $L
timeout 200 python tools/exp_bwd_trace.py 2>&1 | tail -32 >> $L
B200_ATTN_PERSIST=1 timeout 300 python tools/exp_attn2.py --shapes l --skip persist_bwd 2>&1 | tail -2 >> $L
cat $L
