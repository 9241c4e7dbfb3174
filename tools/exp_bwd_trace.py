"""In-kernel timeline of the persistent attention backward (CTA 0), per global tile, in SM cycles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
B, N, H, hd = 128, 256, 32, 160
qkv = (torch.randn(B * N, 3 * H * hd, device="cuda") * 0.5).to(torch.bfloat16)
dout = torch.randn(B * N, H * hd, device="cuda").to(torch.bfloat16)
out, lse = co.attention_fwd_lse(qkv, B, N, H, hd)
for _ in range(2):
    co.attention_bwd_lse(dout, qkv, out, lse, B, N, H, hd, want_colsum=True)
names = ["Y_issued", "Ts_issue", "Tp_issue", "acc_issue", "ts_full", "P_done", "tp_full", "dS_done", "epi_start",
         "epi_end", "X_issued"]
for role, rname in ((0, "dK/dV"), (1, "dQ")):
    tiles = 64
    tr = torch.zeros(tiles * 16, dtype=torch.int64, device="cuda")
    co._C.attention_bwd_set_trace(tr, role)
    co.attention_bwd_lse(dout, qkv, out, lse, B, N, H, hd, want_colsum=True)
    torch.cuda.synchronize()
    co._C.attention_bwd_set_trace(None, 0)
    t = tr.view(tiles, 16).cpu()
    base = int(t[16, 4])
    print(f"role {rname}: cycles relative to ts_full of tile 16 (item 4, tile 0); 4 tiles per item")
    for k in range(16, 28):
        print(k, {names[s]: int(t[k, s]) - base for s in range(11) if int(t[k, s]) != 0})
    per_item = [int(t[k + 4, 4]) - int(t[k, 4]) for k in range(16, 48, 4)]
    print("cycles per item:", per_item)
