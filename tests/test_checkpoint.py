"""Checkpoint contract: per-rank files, four top-level keys, save->resume == uninterrupted run,
and offline consolidation reproduces the unsharded parameters bit-exactly."""
import os

import torch

from dist_worker import launch
from helpers import full_params_of, tiny_cfg
from vit_10b_fsdp_example_b200.consolidate_sharded_ckpts import consolidate_files
from vit_10b_fsdp_example_b200.parallel import FSDPViT


def test_resume_equals_uninterrupted(tmp_path):
    d = str(tmp_path)
    full = launch(2, {"steps": 5}, os.path.join(d, "full.json"))
    part = launch(2, {"steps": 3, "save_at": 3, "save_path": os.path.join(d, "epoch_1_rank_{rank}.ckpt")},
                  os.path.join(d, "part.json"))
    assert part["losses"] == full["losses"][:3]
    for r in range(2):
        ck = torch.load(os.path.join(d, f"epoch_1_rank_{r}.ckpt"), map_location="cpu", weights_only=False)
        assert sorted(ck.keys()) == ["lr_scheduler", "model", "optimizer", "shard_metadata"]
        assert ck["shard_metadata"]["rank"] == r and ck["shard_metadata"]["world_size"] == 2
    rest = launch(2, {"steps": 5, "resume_from": os.path.join(d, "epoch_1_rank_{rank}.ckpt"), "resume_step": 3},
                  os.path.join(d, "rest.json"))
    for a, b in zip(rest["losses"], full["losses"][3:]):
        assert abs(a - b) < 1e-6


def test_no_fsdp_checkpoint_has_no_shard_metadata(tmp_path):
    d = str(tmp_path)
    launch(1, {"steps": 1, "no_fsdp": True, "save_at": 1, "save_path": os.path.join(d, "epoch_1_rank_{rank}.ckpt")},
           os.path.join(d, "r.json"))
    ck = torch.load(os.path.join(d, "epoch_1_rank_0.ckpt"), map_location="cpu", weights_only=False)
    assert ck["shard_metadata"] is None


def test_consolidation_matches_unsharded(tmp_path):
    d = str(tmp_path)
    for flatten in (False, True):
        prefix = os.path.join(d, f"f{int(flatten)}_epoch_1")
        launch(4, {"steps": 0, "flatten": flatten, "seed": 5, "dump_state": prefix + "_rank_{rank}.ckpt"},
               os.path.join(d, "r.json"))
        full = consolidate_files(prefix, save_path=prefix + "_full.pth")
        cfg = tiny_cfg()
        ref = full_params_of(FSDPViT(cfg, dtype=torch.float32, seed=5))
        P = cfg.patch_size
        assert full["pos_embed"].shape == (1, cfg.num_patches, cfg.embed_dim)
        assert full["patch_embed.proj.weight"].shape == (cfg.embed_dim, 3, P, P)
        for k, v in ref.items():
            got = full[k]
            if k == "patch_embed.proj.weight":
                v = v[:, : cfg.patch_k]
            assert torch.equal(got.reshape(-1), v.reshape(-1)), k
        saved = torch.load(prefix + "_full.pth", weights_only=False)
        assert set(saved["model"].keys()) == set(full.keys())


def test_consolidated_checkpoint_round_trip(tmp_path):
    """Sharded files written at world size 4 -> offline consolidation -> (a) a plain nn.Module ViT loads the result
    with strict=True and computes the same logits as the engine, (b) the engine itself restarts from it at a
    *different* world size (1), which the per-rank shard files alone cannot do."""
    from vit_10b_fsdp_example_b200.models.plain import PlainViT

    d = str(tmp_path)
    prefix = os.path.join(d, "epoch_1")
    launch(4, {"steps": 0, "seed": 7, "dump_state": prefix + "_rank_{rank}.ckpt"}, os.path.join(d, "r.json"))
    consolidate_files(prefix, save_path=prefix + "_full.pth")
    cfg = tiny_cfg()
    ref = FSDPViT(cfg, dtype=torch.float32, seed=7)          # what the 4 ranks jointly hold
    other = FSDPViT(cfg, dtype=torch.float32, seed=99)       # different init, then restored from the full file
    full = torch.load(prefix + "_full.pth", weights_only=False)["model"]
    other.load_full_state_dict(full)
    a, b = full_params_of(ref), full_params_of(other)
    for k in a:
        assert torch.equal(a[k], b[k]), k

    plain = PlainViT.from_consolidated(prefix + "_full.pth", cfg).eval()
    images = torch.randn(3, 3, cfg.image_size, cfg.image_size)
    with torch.no_grad():
        want = plain(images)
    got = ref.eval()(images)
    assert torch.allclose(got, want, atol=1e-4), (got - want).abs().max()


def test_consolidated_checkpoints_of_other_stacks_load_after_key_normalisation(tmp_path):
    """Migration path for users of the reference: a consolidated state_dict whose names still carry wrapper prefixes
    (torch_xla FSDP's `_fsdp_wrapped_module.` / `_fpw_module.`, DDP's `module.`) and timm's 4-D patch-embedding
    weight loads into the engine."""
    from vit_10b_fsdp_example_b200.utils.checkpoint import normalize_full_state_dict_keys

    cfg = tiny_cfg()
    src = FSDPViT(cfg, dtype=torch.float32, seed=3)
    full = full_params_of(src)
    foreign = {}
    for k, v in full.items():
        if k.startswith("blocks."):
            i, rest = k.split(".", 2)[1:]
            k2 = f"module._fsdp_wrapped_module._fpw_module.blocks.{i}._fsdp_wrapped_module._fpw_module.{rest}"
        else:
            k2 = "module._fsdp_wrapped_module._fpw_module." + k
        if k == "patch_embed.proj.weight":  # timm stores the conv kernel as [D, 3, P, P]
            v = v[:, : 3 * cfg.patch_size ** 2].reshape(cfg.embed_dim, 3, cfg.patch_size, cfg.patch_size)
        foreign[k2] = v.clone()
    clean = normalize_full_state_dict_keys(foreign)
    assert set(clean) == set(full)
    dst = FSDPViT(cfg, dtype=torch.float32, seed=11)
    dst.load_full_state_dict(clean)
    got = full_params_of(dst)
    for k in full:
        assert torch.equal(got[k], full[k]), k
    # already-clean names pass through; colliding names are an error, not a silent overwrite
    assert set(normalize_full_state_dict_keys(full)) == set(full)
    import pytest
    with pytest.raises(KeyError):
        normalize_full_state_dict_keys({"module.head.bias": 1, "head.bias": 2})
