"""Process launch + runtime helpers: one process per GPU over ``torch.distributed``.

Replaces ``xmp.spawn`` / ``xla_dist`` and the ``xm.*`` runtime helpers the reference uses
(run_vit_training.py:364, README.md:99-118; xm.master_print/rendezvous/mesh_reduce/get_memory_info/
add_step_closure at :203-213,224,289):

  * launched by ``torchrun`` (RANK/WORLD_SIZE/LOCAL_RANK in the env) -> join that job;
  * launched as a plain script -> ``mp.spawn`` one process per visible GPU (or ``--nproc`` on CPU/gloo).
"""
from __future__ import annotations

import os
import socket
from typing import Any, Callable, List

import torch
import torch.distributed as dist


class Runtime:
    """Rank / device / host-collective helpers for one process."""

    def __init__(self, rank: int, world: int, local_rank: int, device: torch.device):
        self.rank, self.world, self.local_rank, self.device = rank, world, local_rank, device
        self._closures: List = []

    # xm.master_print
    def master_print(self, *args, **kwargs) -> None:
        if self.rank == 0:
            print(*args, **kwargs, flush=True)

    # xm.rendezvous(tag): host-side named barrier
    def rendezvous(self, tag: str = "") -> None:
        if self.world > 1:
            dist.barrier()

    # xm.mesh_reduce(tag, value, reduce_fn): host-side gather of python scalars, then reduce_fn(list)
    def mesh_reduce(self, tag: str, value: Any, reduce_fn: Callable[[list], Any]) -> Any:
        if self.world == 1:
            return reduce_fn([value])
        values = [None] * self.world
        dist.all_gather_object(values, value)
        return reduce_fn(values)

    # xm.get_memory_info(device)
    def get_memory_info(self) -> dict:
        if self.device.type == "cuda":
            free, total = torch.cuda.mem_get_info(self.device)
            return {"kb_free": free // 1024, "kb_total": total // 1024,
                    "kb_peak_allocated": torch.cuda.max_memory_allocated(self.device) // 1024}
        try:
            import psutil

            vm = psutil.virtual_memory()
            return {"kb_free": vm.available // 1024, "kb_total": vm.total // 1024}
        except Exception:
            return {"kb_free": 0, "kb_total": 0}

    # xm.add_step_closure(fn, args): run after the step's device work has been enqueued
    def add_step_closure(self, fn: Callable, args=()) -> None:
        self._closures.append((fn, args))

    def run_step_closures(self) -> None:
        closures, self._closures = self._closures, []
        for fn, args in closures:
            fn(*args)


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def pick_device(cfg, local_rank: int) -> torch.device:
    want = getattr(cfg, "device", "auto")
    if want == "cpu" or (want == "auto" and not torch.cuda.is_available()):
        return torch.device("cpu")
    torch.cuda.set_device(local_rank)
    return torch.device("cuda", local_rank)


def init_runtime(cfg, rank: int, world: int, local_rank: int) -> Runtime:
    device = pick_device(cfg, local_rank)
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if device.type == "cuda" else "gloo"
        kwargs = {}
        if device.type == "cuda":
            kwargs["device_id"] = device
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return Runtime(rank, world, local_rank, device)


def _worker(local_rank: int, fn, cfg, world: int, port: int) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(local_rank)
    os.environ["LOCAL_RANK"] = str(local_rank)
    os.environ["WORLD_SIZE"] = str(world)
    rt = init_runtime(cfg, local_rank, world, local_rank)
    try:
        fn(rt, cfg)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def launch(fn: Callable[[Runtime, Any], None], cfg) -> None:
    """Run ``fn(runtime, cfg)`` in every process of the job."""
    # must be in the environment before the first CUDA allocation of this process and of spawned workers: the
    # memory-aware activation policy fills HBM to within a few GB, which needs an allocator that does not fragment
    os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:  # torchrun / torch.distributed.run
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        local_rank = int(os.environ.get("LOCAL_RANK", rank))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        rt = init_runtime(cfg, rank, world, local_rank)
        try:
            fn(rt, cfg)
        finally:
            if dist.is_initialized():
                dist.destroy_process_group()
        return
    nproc = getattr(cfg, "nproc", 0)
    use_cuda = getattr(cfg, "device", "auto") != "cpu" and torch.cuda.is_available()
    if nproc <= 0:
        nproc = torch.cuda.device_count() if use_cuda else 1
    if nproc == 1:
        rt = init_runtime(cfg, 0, 1, 0)
        fn(rt, cfg)
        return
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(fn, cfg, nproc, _free_port()), nprocs=nproc, join=True)
