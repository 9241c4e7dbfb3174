import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for (B, N, H, hd) in [(128, 256, 32, 160), (128, 196, 16, 64)]:
    D = H * hd
    qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
    for fused in (True, False):
        co.FUSED_ATTENTION = fused
        t_nop = timeit(lambda: co.attention_fwd(qkv, B, N, H, hd, need_p=False))
        t_p = timeit(lambda: co.attention_fwd(qkv, B, N, H, hd, need_p=True))
        print(f"B{B} N{N} H{H} hd{hd} fused={fused}: fwd {t_nop:.0f} us, fwd+P {t_p:.0f} us")
