"""Launches the fused attention forward on the ViT-L and ViT-10B shapes (for an ncu capture)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import native
C = native.load()
for (B, N, H, hd) in ((128, 196, 16, 64), (128, 256, 32, 160)):
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        C.attention_fwd(qkv, out, None, None, B, N, H, hd)
torch.cuda.synchronize()
