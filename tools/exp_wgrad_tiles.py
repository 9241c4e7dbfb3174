"""wgrad tail-wave experiment: the weight-gradient GEMMs have few output tiles (proj 400, qkv 1200 tiles of 256x256 over 74
CTA pairs -> 5.4 / 16.2 waves).  Does a 256x128 tile (twice the tiles, finer tail) pay?  Interleaved A/B, one process."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co


def t(fn, n=6, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


T, D = 32768, 5120
x = (torch.randn(T, D, device="cuda") * 0.5).to(torch.bfloat16)
res = {}
for name, N in (("proj", D), ("qkv", 3 * D), ("fc1", 4 * D)):
    dy = torch.randn(T, N, device="cuda").to(torch.bfloat16)
    fl = 2.0 * T * N * D
    a = co.linear_wgrad(dy, x, block_n=256)
    b = co.linear_wgrad(dy, x, block_n=128)
    res[f"{name}_equal"] = bool(torch.equal(a, b))
    r256, r128 = [], []
    for _ in range(4):
        r256.append(round(fl / t(lambda: co.linear_wgrad(dy, x, block_n=256)) / 1e9, 1))
        r128.append(round(fl / t(lambda: co.linear_wgrad(dy, x, block_n=128)) / 1e9, 1))
    res[f"{name}_256"], res[f"{name}_128"] = r256, r128
    del dy, a, b
print(json.dumps(res))
