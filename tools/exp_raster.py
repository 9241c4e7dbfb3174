"""Raster / L2-hint experiment for the tcgen05 GEMM: time (CUDA events) for each setting; run under
`ncu --metrics dram__bytes_read.sum,...` to get DRAM traffic.  Settings come from env vars so each setting is its own
process:  B200_GEMM_GROUP_N, B200_GEMM_HINT_A, B200_GEMM_HINT_B."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co  # noqa: E402

T = 32768
shape = os.environ.get("EXP_SHAPE", "qkv")
K, N = {"qkv": (5120, 15360), "fc2": (20480, 5120), "fc1": (5120, 20480)}[shape]
x = torch.randn(T, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
iters = int(os.environ.get("EXP_ITERS", "10"))
for _ in range(3):
    co.linear_fwd(x, w)
torch.cuda.synchronize()
evs = []
for _ in range(iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    co.linear_fwd(x, w)
    e.record()
    evs.append((s, e))
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in evs)
print(json.dumps({"shape": shape, "group_n": os.environ.get("B200_GEMM_GROUP_N"), "hint_a": os.environ.get("B200_GEMM_HINT_A"),
                  "hint_b": os.environ.get("B200_GEMM_HINT_B"), "ms": ts[len(ts) // 2], "tflops": 2.0 * T * N * K / ts[len(ts) // 2] / 1e9}))
