import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
T, D = 25088, 1024
# rotate over several buffers so that L2 does not hold the working set (8 x 51 MB > 126 MB)
xs = [torch.randn(T, D, device="cuda").to(torch.bfloat16) for _ in range(8)]
g = torch.ones(D, device="cuda", dtype=torch.bfloat16); b = torch.zeros(D, device="cuda", dtype=torch.bfloat16)
i = [0]
def nxt():
    i[0] = (i[0] + 1) % 8
    return xs[i[0]]
y, mean, rstd = co.ln_fwd(xs[0], g, b, 1e-5)
print("ln_small", os.environ.get("B200_LN_SMALL", "1"), "ln_fwd us", round(timeit(lambda: co.ln_fwd(nxt(), g, b, 1e-5)), 1),
      "ln_bwd us", round(timeit(lambda: co.ln_bwd(nxt(), nxt(), g, mean, rstd, dres=nxt(), want_dxsum=True)), 1))
qkv = torch.randn(T, 3 * D, device="cuda").to(torch.bfloat16)
print("attn fwd us", round(timeit(lambda: co.attention_fwd(qkv, 128, 196, 16, 64)), 1))
out, p = co.attention_fwd(qkv, 128, 196, 16, 64)
print("attn bwd us", round(timeit(lambda: co.attention_bwd(out, qkv, p, 128, 196, 16, 64, want_colsum=True)), 1))
u = torch.randn(T, 4 * D, device="cuda").to(torch.bfloat16)
print("gelu us", round(timeit(lambda: co._C.gelu_fwd(u, u)), 1))
