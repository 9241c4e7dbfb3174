"""Launches every hot kernel once per variant on ViT-10B shapes (for one `ncu --set full` capture -> profiles/r2_ncu.md).

    ncu --set full --clock-control none --import-source on -s <warm-up launches> -o gpurun_out/prof_r2 python tools/prof_kernels.py
The script runs the whole list WARM times un-profiled first (`--warm`), prints the number of launches that took, then
runs the list once more; pass that count to `ncu -s`.
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co

T, D, F = 32768, 5120, 20480
B, N, H, hd = 128, 256, 32, 160
bf = torch.bfloat16
x = (torch.randn(T, D, device="cuda") * 0.5).to(bf)
wq = (torch.randn(3 * D, D, device="cuda") * 0.02).to(bf)
bq = torch.randn(3 * D, device="cuda").to(bf)
w1 = (torch.randn(F, D, device="cuda") * 0.02).to(bf)
b1 = torch.randn(F, device="cuda").to(bf)
w2 = (torch.randn(D, F, device="cuda") * 0.01).to(bf)
b2 = torch.randn(D, device="cuda").to(bf)
g = (torch.randn(T, F, device="cuda") * 0.5).to(bf)
dy = torch.randn(T, D, device="cuda").to(bf)
gam = torch.ones(D, device="cuda", dtype=bf); bet = torch.zeros(D, device="cuda", dtype=bf)
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(bf)
n = 314_639_360 // 8
hi = torch.zeros(n, dtype=bf, device="cuda"); lo = torch.zeros(n, dtype=torch.int16, device="cuda")
m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda"); gr = torch.zeros(n, dtype=torch.float32, device="cuda")
clip = torch.ones(1, device="cuda")
shard = torch.randn(75 * 2 ** 20 // 2, device="cuda").to(bf)
full = torch.empty(2 * shard.numel(), device="cuda", dtype=bf)
chunk = co._C.ag_chunk_bytes()
rows, pre = [], 0
for r in range(2):
    rows.append([r, 0, r * shard.numel() * 2, shard.numel() * 2, pre]); pre += -(-shard.numel() * 2 // chunk)
ag_table = torch.tensor(rows, dtype=torch.int64, device="cuda")


def once():
    co.linear_fwd(x, wq, bq)                                   # gemm<0,0> qkv forward, bias epilogue
    u_g = co.linear_fwd(x, w1, b1, act="gelu", want_preact=True)  # gemm<0,0> fc1 forward: GELU + pre-activation output
    co.linear_fwd(g, w2, b2, residual=x)                        # gemm<0,0> fc2 forward: K = 20480, residual epilogue
    co.linear_dgrad(dy, w2, dgelu_preact=u_g[1], want_colsum=True)  # gemm<0,1> fc2 dgrad: dGELU + column sums
    co.linear_dgrad(g, w1)                                      # gemm<0,1> fc1 dgrad (plain)
    co.linear_wgrad(g, x)                                       # gemm<1,1> fc1 wgrad
    y, mean, rstd = co.ln_fwd(x, gam, bet, 1e-5)                # ln_fwd
    co.ln_bwd(dy, x, gam, mean, rstd, dres=x, want_dxsum=True)  # ln_bwd_stream
    out, lse = co.attention_fwd_lse(qkv, B, N, H, hd)           # attn_fwd_persist
    co.attention_bwd_lse(dy, qkv, out, lse, B, N, H, hd, want_colsum=True)  # attn_delta + attn_bwd_persist x 2
    co.adamw_split(hi, lo, m, v, gr, clip, 1e-3, 0.9, 0.999, 1e-8, 0.1, 1)  # adamw_split
    co._C.p2p_all_gather([shard.data_ptr(), shard.data_ptr()], 0, full, ag_table, pre, 64)  # light all-gather (self-peers)
    co.gelu_fwd(u_g[1])                                         # gelu re-materialisation


n0 = co.launch_count()
once()
per = co.launch_count() - n0
for _ in range(2):
    once()
torch.cuda.synchronize()
print(f"launches per pass (hand-written kernels): {per}", flush=True)
once()
torch.cuda.synchronize()
