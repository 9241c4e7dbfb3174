"""Training driver: the same loop as the reference's ``train`` / ``eval_on_val`` / ``run_logging``
(run_vit_training.py:203-324), on the B200-native engine.
"""
from __future__ import annotations

import json
import os
import pprint
import time

import torch

from .config import ViTConfig
from .data import build_datasets
from .launch import Runtime
from .parallel import FSDPViT, GraphedTrainStep, ShardedAdamW
from .utils import SmoothedValue, get_warmup_cosine_scheduler
from .utils.checkpoint import load_ckpt, save_ckpt


def resolve_dtype(cfg, device: torch.device) -> torch.dtype:
    if cfg.dtype == "auto":
        return torch.bfloat16 if device.type == "cuda" else torch.float32
    return torch.bfloat16 if cfg.dtype == "bf16" else torch.float32


def spans_multiple_hosts() -> bool:
    """True when a torchrun job has more ranks than this host has processes (the reference's pod launch,
    README.md:99-101).  The symmetric-memory kernels need every peer on the same NVSwitch domain."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    return world > local


def resolve_backend(cfg, device: torch.device) -> str:
    if cfg.backend == "auto":
        if device.type != "cuda":
            return "torchdist"
        # one NVSwitch box: hand-written P2P / NVLS kernels; several hosts: NCCL collectives (torch.distributed)
        return "torchdist" if spans_multiple_hosts() else "sm100"
    if cfg.backend == "sm100" and spans_multiple_hosts():
        raise ValueError("--backend sm100 needs all ranks on one NVLink/NVSwitch box; use --backend nccl across hosts")
    return "sm100" if cfg.backend == "sm100" else "torchdist"


def build_fsdp_vit_model(cfg, rt: Runtime) -> FSDPViT:
    """Create the ViT with per-block FSDP units + a root unit and gradient checkpointing
    (reference build_fsdp_vit_model, run_vit_training.py:165-200)."""
    return FSDPViT(
        ViTConfig.from_args(cfg), world=rt.world, rank=rt.rank, device=rt.device, dtype=resolve_dtype(cfg, rt.device),
        reshard_after_forward=cfg.reshard_after_forward, flatten_parameters=cfg.flatten_parameters,
        grad_ckpt=cfg.grad_ckpt, run_without_fsdp=cfg.run_without_fsdp, shard_on_cpu=cfg.shard_on_cpu,
        backend=resolve_backend(cfg, rt.device), seed=cfg.seed, verbose_build=rt.master_print,
        ckpt_keep_blocks=getattr(cfg, "ckpt_keep_blocks", 0) if rt.device.type == "cuda" else
        max(0, getattr(cfg, "ckpt_keep_blocks", 0)),
    )


def run_logging(rt: Runtime, cfg, epoch, step, smoothed_loss, smoothed_time, loss, lr, step_ms, global_batch):
    loss_value = loss.item()
    reduced_loss = rt.mesh_reduce("loss_value", loss_value, sum) / rt.world
    smoothed_loss.update(reduced_loss, batch_size=1)
    ips = global_batch / (step_ms * 1e-3) if step_ms and step_ms > 0 else float("nan")
    rt.master_print(
        f"epoch {epoch} step {(step + 1)}, lr: {lr:.4f}, "
        f"loss: {smoothed_loss.avg:.4f}, "
        f"sec/iter: {smoothed_time.avg:.4f}, "
        f"images/sec: {ips:.1f}, "
        f"GPU memory: {rt.get_memory_info()}"
    )
    if cfg.bench_json and rt.rank == 0:
        with open(cfg.bench_json, "a") as f:
            f.write(json.dumps({"epoch": epoch, "step": step + 1, "lr": lr, "loss": smoothed_loss.avg,
                                "sec_per_iter": smoothed_time.avg, "images_per_sec": ips}) + "\n")


@torch.no_grad()
def eval_on_val(rt: Runtime, val_loader, model: FSDPViT, max_steps: int = 0):
    model.eval()
    local_correct = torch.zeros(1, dtype=torch.long, device=rt.device)
    local_total = 0
    for i, (data, target) in enumerate(val_loader):
        output = model(data)
        pred = output.argmax(dim=-1)
        local_correct.add_(pred.eq(target.view_as(pred)).sum())
        local_total += target.size(0)
        if max_steps and i + 1 >= max_steps:
            break
    correct = rt.mesh_reduce("local_correct", local_correct.item(), sum)
    total = rt.mesh_reduce("local_total", local_total, sum)
    accuracy = correct / max(total, 1)
    return accuracy, correct, total


def train(rt: Runtime, cfg):
    batch_size = cfg.batch_size
    num_epochs = cfg.num_epochs
    device = rt.device
    rank = rt.local_rank  # checkpoint files are keyed by the host-local ordinal (reference :220,247,298)

    # build datasets
    train_dataset, train_loader, train_sampler, _, val_loader, _ = build_datasets(
        cfg, device, rt.world, rt.rank, log=rt.master_print)
    rt.rendezvous("loaded dataset")
    rt.master_print(f"\n=== dataset ===\n{pprint.pformat(train_dataset)}\n")

    # build model (loss is fused into the model's step: cross-entropy fwd+bwd kernel)
    model = build_fsdp_vit_model(cfg, rt)
    rt.rendezvous("loaded model")
    rt.master_print(f"\n=== model ===\n{pprint.pformat(model)}\n")

    parameters = list(model.parameters())
    rt.master_print(f"per-GPU (sharded) parameter num: {sum(p.numel() for p in parameters)}")

    # build optimizer and scheduler
    # without gradient clipping the AdamW update is fused into each unit's reduce-scatter kernel
    use_graph = bool(getattr(cfg, "cuda_graph", False)) and device.type == "cuda"
    optimizer = ShardedAdamW(model, lr=cfg.lr, weight_decay=cfg.weight_decay,
                             fuse_into_reduce_scatter=cfg.clip_grad_norm <= 0 and not use_graph)
    graphed = GraphedTrainStep(model, optimizer, cfg.clip_grad_norm) if use_graph else None
    lr_scheduler = get_warmup_cosine_scheduler(
        optimizer, warmup_iteration=cfg.warmup_steps, max_iteration=len(train_dataset) // batch_size * num_epochs)
    rt.rendezvous("loaded optimizer")
    rt.master_print(f"\n=== optimizer ===\n{pprint.pformat(optimizer)}\n")

    if getattr(cfg, "init_from_full_ckpt", ""):
        full = torch.load(cfg.init_from_full_ckpt, map_location="cpu", weights_only=False)
        model.load_full_state_dict(full.get("model", full))
        rt.master_print(f"parameters initialised from the consolidated checkpoint {cfg.init_from_full_ckpt}")
        del full
    # resume (each rank loads its own shard file)
    os.makedirs(cfg.ckpt_dir, exist_ok=True)
    if cfg.resume_epoch > 0:
        ckpt_path = os.path.join(cfg.ckpt_dir, f"epoch_{cfg.resume_epoch}_rank_{rank}.ckpt")
        load_ckpt(ckpt_path, model, optimizer, lr_scheduler)

    smoothed_loss = SmoothedValue(window_size=5)
    smoothed_time = SmoothedValue(window_size=5)
    is_cuda = device.type == "cuda"
    rt.rendezvous("training begins")
    rt.master_print("training begins (kernels are precompiled: no warm-up compilation)")
    for epoch in range(cfg.resume_epoch + 1, num_epochs + 1):
        rt.master_print(f"starting epoch {epoch}")
        time_epoch_b = time_step_b = time.time()
        model.train()
        train_sampler.set_epoch(epoch)
        ev_prev = torch.cuda.Event(enable_timing=True) if is_cuda else None
        if is_cuda:
            ev_prev.record()
        for step, (data, target) in enumerate(train_loader):
            if graphed is not None:
                # whole step (fwd, bwd, collectives, clip, AdamW) replayed as one CUDA graph
                loss = graphed(data, target)
                lr_scheduler.step()
                optimizer.zero_grad(set_to_none=True)
                t_new = time.time()
                time_step_elapsed, time_step_b = t_new - time_step_b, t_new
                smoothed_time.update(time_step_elapsed, batch_size=1)
                is_first_iter = epoch == cfg.resume_epoch + 1 and step == 0
                if is_first_iter or (step + 1) % cfg.log_step_interval == 0:
                    lr = optimizer.param_groups[0]["lr"]
                    ev_now = torch.cuda.Event(enable_timing=True)
                    ev_now.record()
                    ev_now.synchronize()
                    span = 1 if is_first_iter else cfg.log_step_interval
                    step_ms = rt.mesh_reduce("step_ms", ev_prev.elapsed_time(ev_now) / span, max)
                    ev_prev = ev_now
                    run_logging(rt, cfg, epoch, step, smoothed_loss, smoothed_time, loss, lr, step_ms, batch_size)
                if cfg.max_steps and step + 1 >= cfg.max_steps:
                    break
                continue
            # 1+2. forward, loss, backward (explicit hand-written backward; gradients end up reduce-scattered)
            loss = model.forward_backward(data, target)
            if not cfg.run_without_fsdp:
                # clip on the norm of the FULL gradient (reference :266-270)
                if cfg.clip_grad_norm > 0:
                    model.clip_grad_norm_(cfg.clip_grad_norm)
            else:
                # DDP baseline: gradients were all-reduced inside forward_backward (xm.reduce_gradients, :273)
                if cfg.clip_grad_norm > 0:
                    model.clip_grad_norm_(cfg.clip_grad_norm)

            # 3. parameter update
            optimizer.step()
            lr_scheduler.step()
            optimizer.zero_grad(set_to_none=True)

            # 4. logging
            t_new = time.time()
            time_step_elapsed, time_step_b = t_new - time_step_b, t_new
            smoothed_time.update(time_step_elapsed, batch_size=1)
            is_first_iter = epoch == cfg.resume_epoch + 1 and step == 0
            if is_first_iter or (step + 1) % cfg.log_step_interval == 0:
                lr = optimizer.param_groups[0]["lr"]
                step_ms = None
                if is_cuda:
                    ev_now = torch.cuda.Event(enable_timing=True)
                    ev_now.record()
                    ev_now.synchronize()
                    span = 1 if is_first_iter else cfg.log_step_interval
                    step_ms = ev_prev.elapsed_time(ev_now) / span
                    # device time of a step is the max over ranks
                    step_ms = rt.mesh_reduce("step_ms", step_ms, max)
                    ev_prev = ev_now
                else:
                    step_ms = smoothed_time.avg * 1e3
                rt.add_step_closure(run_logging, args=(rt, cfg, epoch, step, smoothed_loss, smoothed_time, loss, lr,
                                                       step_ms, batch_size))
            rt.run_step_closures()
            if cfg.max_steps and step + 1 >= cfg.max_steps:
                break

        time_epoch_elapsed = time.time() - time_epoch_b
        rt.master_print(f"epoch {epoch} done ({time_epoch_elapsed:.2f} sec)")

        # save checkpoint (every rank writes its own shards)
        if epoch % cfg.ckpt_epoch_interval == 0 or epoch == num_epochs:
            ckpt_path = os.path.join(cfg.ckpt_dir, f"epoch_{epoch}_rank_{rank}.ckpt")
            save_ckpt(ckpt_path, model, optimizer, lr_scheduler, master_only=False, rank=rt.rank, barrier=rt.rendezvous)
        # evaluate on val
        if epoch % cfg.test_epoch_interval == 0 or epoch == num_epochs:
            accuracy, _, _ = eval_on_val(rt, val_loader, model, max_steps=cfg.max_steps)
            rt.master_print(f"accuracy on val: {accuracy:.4f}")
    return model, optimizer, lr_scheduler


def main(rt: Runtime, cfg):
    rt.master_print(f"\n=== cfg ===\n{pprint.pformat(cfg)}\n")
    train(rt, cfg)
    rt.master_print("training completed")
