"""Sustained (power-capped) throughput of our GEMM vs cuBLAS: run each for ~3 s back to back and report TFLOP/s,
median SM clock and power.  This is the regime a ViT-10B training step lives in."""
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co  # noqa: E402

T = 32768
D = 5120


def smi():
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip().split(",")
    return float(out[0]), float(out[1])


def sustained(fn, flops, seconds=3.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.time()
    n = 0
    clocks, powers = [], []
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        c, p = smi()
        clocks.append(c)
        powers.append(p)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    clocks.sort()
    return {"ms": ms, "tflops": flops / ms / 1e9, "sm_mhz": clocks[len(clocks) // 2], "power_w": sum(powers) / len(powers)}


cases = []
x = torch.randn(T, D, device="cuda").to(torch.bfloat16)
x4 = torch.randn(T, 4 * D, device="cuda").to(torch.bfloat16)
wq = (torch.randn(3 * D, D, device="cuda") * 0.02).to(torch.bfloat16)
w2 = (torch.randn(D, 4 * D, device="cuda") * 0.02).to(torch.bfloat16)
dy = torch.randn(T, D, device="cuda").to(torch.bfloat16)
cases.append(("qkv_fwd", lambda: co.linear_fwd(x, wq), lambda: torch.nn.functional.linear(x, wq), 2.0 * T * D * 3 * D))
cases.append(("fc2_fwd(K=20480)", lambda: co.linear_fwd(x4, w2), lambda: torch.nn.functional.linear(x4, w2), 2.0 * T * D * 4 * D))
cases.append(("fc2_dgrad", lambda: co.linear_dgrad(dy, w2), lambda: dy @ w2, 2.0 * T * D * 4 * D))
cases.append(("fc2_wgrad(K=32768)", lambda: co.linear_wgrad(dy, x4), lambda: dy.t() @ x4, 2.0 * T * D * 4 * D))
for name, ours, ref, fl in cases:
    for impl, fn in (("ours", ours), ("cublas", ref)):
        r = sustained(fn, fl)
        r.update(case=name, impl=impl)
        print(json.dumps(r), flush=True)
        time.sleep(1.0)
