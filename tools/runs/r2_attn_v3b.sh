#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_attn_v3b.log; : > $L
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -8 >> $L
timeout 300 python tools/exp_attn2.py --shapes 10b,l,336 --skip persist_bwd 2>&1 | tail -4 >> $L
cat $L
