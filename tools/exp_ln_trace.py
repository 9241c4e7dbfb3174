"""(1) LayerNorm fwd / bwd timing at the ViT-10B shape; (2) in-kernel timeline of the persistent attention forward
(clock64 stamps of CTA 0, csrc/attention_persist_sm100.cu) printed per work item in SM cycles."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


T, D = 32768, 5120
x = (torch.randn(T, D, device="cuda") + 1.0).to(torch.bfloat16)
dy = torch.randn(T, D, device="cuda").to(torch.bfloat16)
dres = torch.randn(T, D, device="cuda").to(torch.bfloat16)
g = torch.ones(D, device="cuda", dtype=torch.bfloat16); b = torch.zeros(D, device="cuda", dtype=torch.bfloat16)
y, mean, rstd = co.ln_fwd(x, g, b, 1e-5)
res = {"ln_stream_env": os.environ.get("B200_LN_STREAM", "1")}
res["ln_fwd_us"] = round(timeit(lambda: co.ln_fwd(x, g, b, 1e-5)), 1)
res["ln_bwd_res_dxsum_us"] = round(timeit(lambda: co.ln_bwd(dy, x, g, mean, rstd, dres=dres, want_dxsum=True)), 1)
res["ln_bwd_plain_us"] = round(timeit(lambda: co.ln_bwd(dy, x, g, mean, rstd)), 1)
gb = 4 * T * D * 2 / 1e9
res["ln_bwd_res_GBs"] = round(gb / (res["ln_bwd_res_dxsum_us"] * 1e-6), 0)
print(json.dumps(res), flush=True)

if "--trace" in sys.argv:
    B, N, H, hd = 128, 256, 32, 160
    qkv = (torch.randn(B * N, 3 * H * hd, device="cuda") * 0.5).to(torch.bfloat16)
    out = torch.empty(B * N, H * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H, N, device="cuda")
    for _ in range(2):
        co._C.attention_fwd_persist(qkv, out, lse, B, N, H, hd)
    items = 24
    tr = torch.zeros(items * 16, dtype=torch.int64, device="cuda")
    co._C.attention_set_trace(tr)
    co._C.attention_fwd_persist(qkv, out, lse, B, N, H, hd)
    torch.cuda.synchronize()
    co._C.attention_set_trace(None)
    t = tr.view(items, 16).cpu()
    names = {0: "qk_issued", 1: "qk_full", 2: "S_issue", 3: "acc_empty", 4: "pv0", 5: "pv1", 6: "pv2", 7: "pv3",
             8: "s_full", 9: "pass1", 10: "pass2", 11: "acc_full", 12: "epi_done"}
    base = int(t[4, 8])
    print("attention persist fwd, CTA 0, cycles relative to s_full of item 4")
    for i in range(4, 12):
        row = {names[k]: int(t[i, k]) - base for k in sorted(names) if int(t[i, k]) != 0}
        print(i, row)
    per_item = [(int(t[i + 1, 8]) - int(t[i, 8])) for i in range(4, 20)]
    print("cycles between consecutive s_full:", per_item)
