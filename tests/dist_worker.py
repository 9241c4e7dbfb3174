"""Worker used by the multi-process (gloo) tests: trains a tiny ViT and dumps the loss trajectory."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def run(rank, world, port, opts, out_path):
    from helpers import tiny_cfg
    from vit_10b_fsdp_example_b200.parallel import FSDPViT, ShardedAdamW
    from vit_10b_fsdp_example_b200.utils import get_warmup_cosine_scheduler
    from vit_10b_fsdp_example_b200.utils.checkpoint import load_ckpt, save_ckpt

    torch.set_num_threads(1)
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    if opts.get("poison"):  # NaN-fill every gathered-parameter buffer on release: catches use-after-reshard
        from vit_10b_fsdp_example_b200.parallel import engine as _engine

        _engine.DEBUG_POISON = True
    cfg = tiny_cfg(**opts.get("model", {}))
    model = FSDPViT(cfg, world=world, rank=rank, dtype=torch.float32,
                    reshard_after_forward=opts.get("reshard", True), flatten_parameters=opts.get("flatten", False),
                    grad_ckpt=opts.get("grad_ckpt", True), run_without_fsdp=opts.get("no_fsdp", False),
                    shard_on_cpu=opts.get("shard_on_cpu", False), seed=opts.get("seed", 0),
                    ckpt_keep_blocks=opts.get("keep_blocks", 0))
    opt = ShardedAdamW(model, lr=opts.get("lr", 1e-2), weight_decay=0.1)
    sched = get_warmup_cosine_scheduler(opt, 2, 100)
    global_batch = opts.get("global_batch", 8)
    local = global_batch // world
    g = torch.Generator().manual_seed(1234)
    steps = opts.get("steps", 4)
    images = torch.randn(8, global_batch, 3, cfg.image_size, cfg.image_size, generator=g)
    targets = torch.randint(0, cfg.num_classes, (8, global_batch), generator=g)
    start = 0
    if opts.get("resume_from"):
        load_ckpt(opts["resume_from"].format(rank=rank), model, opt, sched)
        start = opts["resume_step"]
    losses, norms = [], []
    for s in range(start, steps):
        img = images[s, rank * local:(rank + 1) * local]
        tgt = targets[s, rank * local:(rank + 1) * local]
        loss = model.forward_backward(img, tgt)
        norm = model.clip_grad_norm_(opts.get("clip", 1.0))
        opt.step()
        sched.step()
        opt.zero_grad()
        lv = torch.tensor([loss.item()])
        if world > 1:
            dist.all_reduce(lv)
        losses.append(lv.item() / world)
        norms.append(norm.item())
        if opts.get("save_at") == s + 1:
            save_ckpt(opts["save_path"].format(rank=rank), model, opt, sched, master_only=False, rank=rank)
    result = {"losses": losses, "norms": norms, "sharded": model.num_sharded_parameters()}
    if opts.get("dump_state"):
        torch.save({"model": model.state_dict(), "shard_metadata": model.get_shard_metadata(),
                    "optimizer": opt.state_dict(), "lr_scheduler": sched.state_dict()},
                   opts["dump_state"].format(rank=rank))
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(result, f)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def launch(world, opts, out_path):
    from helpers import free_port

    if world == 1:
        run(0, 1, 0, opts, out_path)
    else:
        import torch.multiprocessing as mp

        mp.spawn(run, args=(world, free_port(), opts, out_path), nprocs=world, join=True)
    with open(out_path) as f:
        return json.load(f)
