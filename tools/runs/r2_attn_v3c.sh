#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_attn_v3c.log; : > $L
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -m gpu 2>&1 | tail -5 >> $L
timeout 300 python tools/exp_attn2.py --shapes 10b,l,336 --skip persist_bwd 2>&1 | tail -4 >> $L
B200_ATTN_PERSIST=1 timeout 300 python tools/exp_attn2.py --shapes l --skip persist_bwd 2>&1 | tail -2 >> $L
timeout 200 python tools/exp_bwd_trace.py 2>&1 | grep -E "cycles per item|^19 |^23 " >> $L
cat $L
