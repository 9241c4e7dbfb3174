"""Host-side tables of the symmetric-memory collectives (parallel/backends.py), checked on the CPU with a stand-in
backend object: every byte of every rank's shard must land exactly once in the gathered buffer (pull-kernel segment
table and copy-engine copy lists), sources must be walked in a rank-staggered order, and the reduce-scatter table must
cover this rank's slices exactly once with chunk prefixes that match the kernel's chunking."""
import types

import pytest
import torch

from vit_10b_fsdp_example_b200.config import ViTConfig
from vit_10b_fsdp_example_b200.models import vit
from vit_10b_fsdp_example_b200.parallel import backends
from vit_10b_fsdp_example_b200.parallel.layout import UnitLayout

AG_CHUNK = 16384
RS_VECS = 512


def _fake(world, rank):
    be = types.SimpleNamespace(world=world, rank=rank, _seg_cache={}, device="cpu",
                               _C=types.SimpleNamespace(ag_chunk_bytes=lambda: AG_CHUNK, rs_chunk_vecs=lambda: RS_VECS))
    be._lay_key = backends.Sm100Backend._lay_key
    return be


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("flatten", [False, True])
def test_all_gather_tables_cover_every_byte_once(world, flatten):
    cfg = ViTConfig(embed_dim=640, num_heads=4, num_blocks=1)
    lay = UnitLayout.build("blocks.0", vit.block_param_specs(cfg), world, flatten)
    es = 2
    first_sources = set()
    for rank in range(world):
        be = _fake(world, rank)
        table, chunks = backends.Sm100Backend._ag_table(be, lay, es)
        cover = torch.zeros(lay.full_numel * es, dtype=torch.int32)
        prefix = 0
        for r, src_off, dst_off, n, pf in table.tolist():
            assert pf == prefix and n > 0 and n % 16 == 0
            prefix += -(-n // AG_CHUNK)
            cover[dst_off: dst_off + n] += 1
            # the byte at shard offset o of rank r belongs at (group full offset + r * shard_len) * es + (o - group shard offset)
            g = next(g for g in lay.groups if g.shard_offset * es <= src_off < (g.shard_offset + g.shard_len) * es)
            assert dst_off - (g.full_offset + r * g.shard_len) * es == src_off - g.shard_offset * es
        assert prefix == chunks
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        first_sources.add(int(table[0, 0]))
        # copy-engine lists: one copy per (group, source), own slice last
        shard = torch.zeros(lay.shard_numel, dtype=torch.bfloat16)
        full = torch.zeros(lay.full_numel, dtype=torch.bfloat16)
        be._peer = {shard.data_ptr(): [1_000_000_000 * (r + 1) for r in range(world)]}
        src, dst, nb = backends.Sm100Backend._ag_copies(be, lay, shard, full, ())
        assert len(src) == len(lay.groups) * world
        cover.zero_()
        for s_ptr, d_ptr, n in zip(src, dst, nb):
            off = d_ptr - full.data_ptr()
            cover[off: off + n] += 1
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        assert src[-1] // 1_000_000_000 - 1 == rank, "own slice is copied last"
    assert len(first_sources) == world, "ranks must start at different sources"


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("esize", [2, 4])
def test_reduce_scatter_table(world, esize):
    cfg = ViTConfig(embed_dim=640, num_heads=4, num_blocks=1)
    lay = UnitLayout.build("blocks.0", vit.block_param_specs(cfg), world, False)
    per_chunk = RS_VECS * (16 // esize)
    for rank in range(world):
        be = _fake(world, rank)
        table, chunks = backends.Sm100Backend._rs_table(be, lay, esize)
        cover = torch.zeros(lay.shard_numel, dtype=torch.int32)
        prefix = 0
        for (foff_b, soff, n, pf), (foff, soff2, n2) in zip(table.tolist(), lay.scatter_segments(rank)):
            assert (foff_b, soff, n) == (foff * esize, soff2, n2) and pf == prefix
            assert n % (16 // esize) == 0, "segments are whole 16-byte vectors"
            prefix += -(-n // per_chunk)
            cover[soff: soff + n] += 1
        assert prefix == chunks
        assert int(cover.min()) == 1 and int(cover.max()) == 1


def test_activation_policy_has_no_budget_for_p_under_fused_attention(monkeypatch):
    from helpers import tiny_cfg
    from vit_10b_fsdp_example_b200.ops import torch_ops
    from vit_10b_fsdp_example_b200.parallel import FSDPViT

    model = FSDPViT(tiny_cfg(), dtype=torch.float32, seed=1)
    sizes = dict(model.extra_bytes_per_block(4))
    assert sizes["P"] > 0 and sizes["h"] > 0 and sizes["g"] > 0
    monkeypatch.setattr(torch_ops, "FLASH_ATTENTION", True)
    assert dict(model.extra_bytes_per_block(4))["P"] == 0


def test_keep_policy_is_sized_after_the_first_step_of_this_process_even_when_resumed(monkeypatch):
    """A resumed run restores step_count > 0 (it seeds the dropout masks) but has not seen its transient memory peak
    yet: the automatic keep policy must still run its first step fully checkpointed and size itself afterwards."""
    from helpers import tiny_cfg
    from vit_10b_fsdp_example_b200.parallel import FSDPViT

    cfg = tiny_cfg()
    model = FSDPViT(cfg, dtype=torch.float32, seed=1, ckpt_keep_blocks=-1)
    model.load_state_dict(model.state_dict(), shard_metadata=dict(model.get_shard_metadata(), step_count=7))
    assert model.step_count == 7 and model.keep_blocks == -1
    calls = []
    monkeypatch.setattr(model, "_auto_keep_blocks", lambda batch: calls.append(batch) or 1)
    x = torch.zeros(2, 3, cfg.image_size, cfg.image_size)
    y = torch.zeros(2, dtype=torch.long)
    model.forward_backward(x, y)
    assert calls == [] and model.keep_blocks == -1 and model.step_count == 8
    model.forward_backward(x, y)
    assert calls == [2] and model.keep_blocks == 1


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("which", ["vit10b_root", "vitl_block", "vitl_root"])
def test_tables_on_the_real_units_with_awkward_shapes(world, which):
    """Real unit shapes: the ViT-10B root unit (patch embedding zero-padded to the TMA-legal K, position embedding,
    1000-class head: 125 rows and a 125-element bias slice per rank at W = 8) and ViT-L.  Every segment must stay a whole
    number of 16-byte vectors and cover its buffer exactly once, for the gather and for the reduce-scatter."""
    cfg = ViTConfig() if which == "vit10b_root" else ViTConfig(embed_dim=1024, num_heads=16, num_blocks=24, patch_size=16)
    specs = vit.root_param_specs(cfg) if which.endswith("root") else vit.block_param_specs(cfg)
    lay = UnitLayout.build("u", specs, world, False)
    es = 2
    assert lay.shard_numel * world == lay.full_numel
    for rank in (0, world - 1):
        be = _fake(world, rank)
        table, chunks = backends.Sm100Backend._ag_table(be, lay, es)
        cover = torch.zeros(lay.full_numel * es, dtype=torch.int8)
        for r, src_off, dst_off, n, pf in table.tolist():
            assert n % 16 == 0 and src_off % 16 == 0 and dst_off % 16 == 0
            cover[dst_off: dst_off + n] += 1
        assert int(cover.min()) == 1 and int(cover.max()) == 1
        rs, rchunks = backends.Sm100Backend._rs_table(be, lay, es)
        scover = torch.zeros(lay.shard_numel, dtype=torch.int8)
        for foff_b, soff, n, pf in rs.tolist():
            assert foff_b % 16 == 0 and (soff * 4) % 16 == 0 and (n * es) % 16 == 0
            scover[soff: soff + n] += 1
        assert int(scover.min()) == 1 and int(scover.max()) == 1
