#!/bin/bash
# 1 GPU: whole GPU test-suite + reduced-depth bench of the 224 / 336 configs, both arms, + CLI run
mkdir -p gpurun_out; L=gpurun_out/r2_smoke1.log; : > $L
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 >> $L
echo "== smoke()" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $L
for m in vit10b vit10b_336; do
  echo "== bench $m 4 blocks ours" >> $L
  timeout 600 python bench.py --model $m --num_blocks 4 --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-1500 >> $L
  echo "== bench $m 4 blocks reference" >> $L
  timeout 600 python bench.py --model $m --num_blocks 4 --steps 4 --warmup 3 --impl reference 2>&1 | tail -1 | cut -c1-900 >> $L
done
echo "== CLI (1 GPU, ViT-10B dims, 2 blocks, 3 steps, ckpt + eval)" >> $L
timeout 600 python run_vit_training.py --fake_data --num_blocks 2 --batch_size 64 --num_epochs 1 --max_steps 3 --log_step_interval 1 --ckpt_dir /tmp/ck1 --nproc 1 2>&1 | grep -E "epoch|accuracy|saved|Error|error" | tail -8 >> $L
cat $L
