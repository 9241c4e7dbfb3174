"""xm.* helpers on CUDA + torch.distributed (stock PyTorch only)."""
import os

import torch
import torch.distributed as dist


def xla_device():
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    return torch.device("cuda", local_rank)


def xrt_world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def get_ordinal():
    return dist.get_rank() if dist.is_initialized() else 0


def get_local_ordinal():
    return int(os.environ.get("LOCAL_RANK", 0))


def master_print(*args, **kwargs):
    if get_ordinal() == 0:
        print(*args, **kwargs, flush=True)


def rendezvous(tag):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def mesh_reduce(tag, value, reduce_fn):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return reduce_fn([value])
    values = [None] * dist.get_world_size()
    dist.all_gather_object(values, value)
    return reduce_fn(values)


def add_step_closure(fn, args=()):
    fn(*args)


def reduce_gradients(optimizer):
    world = xrt_world_size()
    if world == 1:
        return
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p.grad is not None:
                dist.all_reduce(p.grad)
                p.grad.div_(world)


def save(obj, path, master_only=True, global_master=False):
    if not master_only or get_ordinal() == 0:
        torch.save(obj, path)
    rendezvous("save")


def get_memory_info(device):
    free, total = torch.cuda.mem_get_info(device)
    return {"kb_free": free // 1024, "kb_total": total // 1024}
