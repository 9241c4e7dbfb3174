"""cProfile of the host side of a ViT-L training step (which Python lines cost the launch-bound configs)."""
import cProfile, io, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_10b_fsdp_example_b200.config import ViTConfig
from vit_10b_fsdp_example_b200.parallel import FSDPViT, ShardedAdamW

dev = torch.device("cuda")
cfg = ViTConfig(image_size=224, patch_size=16, embed_dim=1024, num_heads=16, num_blocks=24, mlp_ratio=4.0)
model = FSDPViT(cfg, device=dev, dtype=torch.bfloat16, init_device="cuda")
opt = ShardedAdamW(model, lr=1e-3, weight_decay=0.1)
x = torch.zeros(128, 3, 224, 224, device=dev); y = torch.zeros(128, dtype=torch.long, device=dev)
def step():
    loss = model.forward_backward(x, y); model.clip_grad_norm_(1.0); opt.step(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5): step()
t_host = time.time() - t0
torch.cuda.synchronize()
t_all = time.time() - t0
print(f"host enqueue {t_host/5*1e3:.1f} ms/step, wall {t_all/5*1e3:.1f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
