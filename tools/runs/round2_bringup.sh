#!/bin/bash
# First GPU call of round 2: bring up everything that was written after the round-1 GPU budget ran out.
#   gpurun --timeout 1500 -- 'bash tools/runs/round2_bringup.sh'          (1 GPU)
mkdir -p gpurun_out; L=gpurun_out/bringup.log; : > $L
echo "== gated tests (long-sequence attention kernels, engine path through the fused attention backward)" >> $L
B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_engine.py -q -m gpu 2>&1 | tail -15 >> $L
echo "== persistent forward vs one-shot / un-fused forward (ViT-10B shape)" >> $L
timeout 120 python - >> $L 2>&1 <<'PY'
import torch, os, sys
sys.path.insert(0, os.getcwd())
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
B, N, H, hd = 128, 256, 32, 160
qkv = (torch.randn(B * N, 3 * H * hd, device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
def t(fn, n=10):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2] * 1e3
print("persistent us", t(lambda: co._C.attention_fwd_persist(qkv, out, None, B, N, H, hd)))
print("one-shot fused us", t(lambda: co._C.attention_fwd(qkv, out, None, None, B, N, H, hd)))
print("un-fused us", t(lambda: co.attention_fwd(qkv, B, N, H, hd, need_p=True)))
PY
echo "== attention fwd/bwd timings" >> $L
B200_TEST_UNVERIFIED=1 timeout 120 python tools/exp_attn_bwd.py 2>&1 | tail -4 >> $L
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
