#!/bin/bash
# long-sequence path (336 px -> 576 tokens/image) end to end through the engine on a 4-block ViT-10B slice
mkdir -p gpurun_out
timeout 150 python bench.py --model vit10b_336 --num_blocks 4 --local_batch 64 --steps 4 --warmup 3 --no_full_ckpt_probe 2>&1 | tail -1 | tee gpurun_out/vit10b_336_slice.log
