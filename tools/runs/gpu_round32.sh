#!/bin/bash
# A/B on one box: lean-only vs lean + extras (P, LN outputs, gelu(u)) for an 8-block ViT-10B slice
mkdir -p gpurun_out; rm -f gpurun_out/extras_ab.log
for e in 0 1 0 1; do
  B200_CKPT_EXTRAS=$e timeout 120 python bench.py --num_blocks 8 --steps 6 --warmup 3 --no_e2e --no_full_ckpt_probe 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print('extras=$e', r['ms_per_step'], r['config']['activation_ckpt'], r['peak_mem_gb'])" >> gpurun_out/extras_ab.log 2>&1
done
cat gpurun_out/extras_ab.log
