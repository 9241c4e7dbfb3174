"""Per-rank sharded checkpoints (reference: utils.py:24-43, run_vit_training.py:245-248,297-299).

Layout kept compatible with the reference: one ``torch.save`` file per rank per epoch named
``epoch_{E}_rank_{R}.ckpt`` holding exactly four keys -- ``model`` (this rank's shards), ``shard_metadata``
(``None`` when not FSDP), ``optimizer`` and ``lr_scheduler``.  Every rank writes its own file; loading goes
through the host (``map_location='cpu'``).  ``consolidate_sharded_ckpts`` rebuilds a full state_dict.
"""
from __future__ import annotations

import os

import torch


def save_ckpt(ckpt_path: str, model, optimizer, lr_scheduler, master_only: bool = True, rank: int = 0,
              barrier=None) -> None:
    ckpt = {
        "model": model.state_dict(),
        "shard_metadata": model.get_shard_metadata() if getattr(model, "use_fsdp", False) else None,
        "optimizer": optimizer.state_dict(),
        "lr_scheduler": lr_scheduler.state_dict(),
    }
    if not master_only or rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(ckpt_path)), exist_ok=True)
        tmp = ckpt_path + ".tmp"
        torch.save(ckpt, tmp)
        os.replace(tmp, ckpt_path)  # never leave a half-written checkpoint behind
    if barrier is not None:
        barrier()  # xm.save contains a rendezvous
    print(f"checkpoint saved to {ckpt_path}\n", end="")


def load_ckpt(ckpt_path: str, model, optimizer, lr_scheduler) -> None:
    assert os.path.exists(ckpt_path), f"checkpoint {ckpt_path} does not exist"
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    try:
        model.load_state_dict(ckpt["model"], shard_metadata=ckpt.get("shard_metadata"))
    except TypeError:  # a plain nn.Module-style model
        model.load_state_dict(ckpt["model"])
    optimizer.load_state_dict(ckpt["optimizer"])
    if getattr(model, "step_count", None) == 0 and hasattr(optimizer, "state") and hasattr(model, "all_units"):
        # DDP checkpoints carry no shard metadata: the optimizer step count is the number of steps taken
        model.step_count = int(optimizer.state[model.all_units[0].name]["step"])
    lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
    print(f"resumed from checkpoint {ckpt_path}\n", end="")


_WRAPPER_PARTS = ("_fsdp_wrapped_module.", "_fpw_module.", "_checkpoint_wrapped_module.", "_orig_mod.")


def normalize_full_state_dict_keys(state: dict) -> dict:
    """Strip wrapper prefixes from the parameter names of a consolidated checkpoint so that one written by another
    stack loads here: torch_xla FSDP (``_fsdp_wrapped_module.`` / ``_fpw_module.``, what the reference's
    ``consolidate_sharded_ckpts`` may leave behind), PyTorch FSDP / activation-checkpoint wrappers, DDP's leading
    ``module.`` and torch.compile's ``_orig_mod.``.  Names that are already timm-style pass through unchanged."""
    out = {}
    for k, v in state.items():
        name = k
        for part in _WRAPPER_PARTS:
            name = name.replace(part, "")
        while name.startswith("module."):
            name = name[len("module."):]
        if name in out:
            raise KeyError(f"checkpoint keys {k!r} and another entry both normalise to {name!r}")
        out[name] = v
    return out
