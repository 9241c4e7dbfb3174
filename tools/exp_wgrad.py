import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.ops import cuda_ops as co
T, D = 32768, 5120
which = os.environ.get("EXP_CASE", "wgrad_fc2")
dy = torch.randn(T, D, device="cuda").to(torch.bfloat16)
x4 = torch.randn(T, 4 * D, device="cuda").to(torch.bfloat16)
w2 = (torch.randn(D, 4 * D, device="cuda") * 0.02).to(torch.bfloat16)
u = torch.randn(T, 4 * D, device="cuda").to(torch.bfloat16)
fn = {"wgrad_fc2": lambda: co.linear_wgrad(dy, x4), "wgrad_fc1": lambda: co.linear_wgrad(x4, dy),
      "dgrad_fc2_dgelu": lambda: co.linear_dgrad(dy, w2, dgelu_preact=u, want_colsum=True),
      "dgrad_fc2": lambda: co.linear_dgrad(dy, w2)}[which]
for _ in range(4):
    fn()
torch.cuda.synchronize()
evs = []
for _ in range(8):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); fn(); e.record(); evs.append((s, e))
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in evs)
print(which, "ms", ts[len(ts)//2], "TF", 2.0 * T * D * 4 * D / ts[len(ts)//2] / 1e9)
