#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_attn_v3.log; : > $L
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -x -q -m gpu 2>&1 | tail -8 >> $L
timeout 200 python tools/exp_ln_trace.py --trace 2>&1 | tail -13 >> $L
timeout 300 python tools/exp_attn2.py --shapes 10b --skip persist_bwd 2>&1 | tail -3 >> $L
echo "== bench vit10b_336 4 blocks ours" >> $L
timeout 600 python bench.py --model vit10b_336 --num_blocks 4 --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-1500 >> $L
cat $L
