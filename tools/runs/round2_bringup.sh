#!/bin/bash
# First GPU call of round 2: bring up everything that was written after the round-1 GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/runs/round2_bringup.sh'          (1 GPU)
mkdir -p gpurun_out; L=gpurun_out/bringup.log; : > $L
echo "== gated tests (long-sequence attention kernels, engine path through the fused attention backward)" >> $L
B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -q -m gpu 2>&1 | tail -15 >> $L
echo "== attention fwd/bwd timings" >> $L
timeout 120 python tools/exp_attn_bwd.py 2>&1 | tail -4 >> $L
for flash in 0 1; do
  echo "== ViT-L, B200_FUSED_ATTN_BWD=$flash" >> $L
  B200_FUSED_ATTN_BWD=$flash timeout 300 python bench.py --model vitl --steps 20 --warmup 3 --no_e2e 2>&1 | tail -1 | cut -c1-400 >> $L
done
echo "== ViT-10B 336 px slice with the long-sequence kernels" >> $L
for long in 0 1; do
  B200_FUSED_ATTN_BWD=$long B200_FUSED_ATTN_LONG=$long timeout 300 python bench.py --model vit10b_336 --num_blocks 4 --local_batch 64 \
      --steps 4 --warmup 3 --no_e2e --no_full_ckpt_probe 2>&1 | tail -1 | cut -c1-300 >> $L
done
cat $L
