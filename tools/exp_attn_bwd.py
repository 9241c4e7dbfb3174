"""Fused (lse) vs un-fused attention backward, CUDA-event medians."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
def t(fn, n=8):
    for _ in range(2): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2] * 1e3
for (B, N, H, hd) in ((128, 256, 32, 160), (128, 196, 16, 64)):
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    dout = torch.randn(B * N, D, device="cuda").to(torch.bfloat16)
    out, lse = co.attention_fwd_lse(qkv, B, N, H, hd)
    _, p = co.attention_fwd(qkv, B, N, H, hd, need_p=True)
    r = dict(shape=(B, N, H, hd),
             fused_bwd_us=t(lambda: co.attention_bwd_lse(dout, qkv, out, lse, B, N, H, hd, want_colsum=True)),
             fused_bwd_persistent_us=(setattr(co, "ATTN_PERSIST", True),
                                      t(lambda: co.attention_bwd_lse(dout, qkv, out, lse, B, N, H, hd, want_colsum=True)),
                                      setattr(co, "ATTN_PERSIST", False))[1] if os.environ.get("B200_TEST_UNVERIFIED") == "1" else None,
             unfused_bwd_us=t(lambda: co.attention_bwd(dout, qkv, p, B, N, H, hd, want_colsum=True)),
             probs_remat_us=t(lambda: co.attention_probs(qkv, B, N, H, hd)),
             fused_fwd_lse_us=t(lambda: co.attention_fwd_lse(qkv, B, N, H, hd)),
             default_fwd_us=t(lambda: co.attention_fwd(qkv, B, N, H, hd, need_p=False)))
    print(r, flush=True)
