"""Driver for ncu captures of the attention kernels (persistent fwd / bwd at the ViT-10B shape) and LayerNorm bwd.
    ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_persist|attn_bwd_persist|ln_bwd' \
        -s 8 -c 4 -o gpurun_out/prof_attn python tools/prof_attn_ln.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
B, N, H, hd = 64, 256, 32, 160
D = H * hd
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
dout = torch.randn(B * N, D, device="cuda").to(torch.bfloat16)
x = torch.randn(B * N, D, device="cuda").to(torch.bfloat16)
g = torch.ones(D, device="cuda", dtype=torch.bfloat16); b = torch.zeros(D, device="cuda", dtype=torch.bfloat16)
for it in range(3):
    out, lse = co.attention_fwd_lse(qkv, B, N, H, hd)                 # attn_fwd_persist
    dqkv = co.attention_bwd_lse(dout, qkv, out, lse, B, N, H, hd)    # delta, attn_bwd_persist x2
    y, mean, rstd = co.ln_fwd(x, g, b, 1e-5)
    dx, dg, db, dxs = co.ln_bwd(x, x, g, mean, rstd, dres=x, want_dxsum=True)
torch.cuda.synchronize()
