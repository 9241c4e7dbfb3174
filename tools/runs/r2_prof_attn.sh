#!/bin/bash
# 1 GPU: gpu tests added this round + ncu of persistent attention and LN bwd kernels
mkdir -p gpurun_out; L=gpurun_out/r2_prof_attn.log; : > $L
timeout 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_attention.py -x -q -m gpu 2>&1 | tail -8 >> $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_persist|attn_bwd_persist|ln_bwd' -s 8 -c 4 -f -o gpurun_out/prof_attn python tools/prof_attn_ln.py >> $L 2>&1
ls -la gpurun_out/*.ncu-rep >> $L
tail -30 $L
