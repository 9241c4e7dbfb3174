"""Runs each hot kernel a few times on ViT-10B shapes (for ncu captures / roofline fractions)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
T, D = 32768, 5120
x = torch.randn(T, D, device="cuda").to(torch.bfloat16)
wq = (torch.randn(3 * D, D, device="cuda") * 0.02).to(torch.bfloat16)
bq = torch.randn(3 * D, device="cuda").to(torch.bfloat16)
g = torch.ones(D, device="cuda", dtype=torch.bfloat16); b = torch.zeros(D, device="cuda", dtype=torch.bfloat16)
n = 314_639_360
hi = torch.zeros(n, dtype=torch.bfloat16, device="cuda"); lo = torch.zeros(n, dtype=torch.int16, device="cuda")
m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda"); gr = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
clip = torch.ones(1, device="cuda")
for it in range(3):
    qkv = co.linear_fwd(x, wq, bq)
    y, mean, rstd = co.ln_fwd(x, g, b, 1e-5)
    dx, dg, db, dxs = co.ln_bwd(x, x, g, mean, rstd, dres=x, want_dxsum=True)
    co.adamw_split(hi, lo, m, v, gr, clip, 1e-3, 0.9, 0.999, 1e-8, 0.1, it + 1)
    out, p = co.attention_fwd(qkv, 128, 256, 32, 160)
torch.cuda.synchronize()
