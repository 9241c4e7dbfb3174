"""Attention core (batched tcgen05 GEMMs + fused softmax) vs the fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(got, ref, rel=3e-2):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    denom = ref.abs().max().item() + 1e-6
    assert err / denom < rel, f"max abs err {err} vs ref max {denom}"


@pytest.mark.parametrize("B,N,H,hd", [(2, 256, 4, 160), (3, 196, 3, 64), (1, 576, 2, 160)])
def test_attention_fwd_bwd(B, N, H, hd):
    from vit_10b_fsdp_example_b200.ops import cuda_ops as co, torch_ops as to

    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    out, p = co.attention_fwd(qkv, B, N, H, hd)
    outr, pr = to.attention_fwd(qkv.float(), B, N, H, hd)
    _close(p.view(B, H, N, -1)[..., :N], pr)
    _close(out, outr)
    dout = (torch.randn(B * N, D, device="cuda")).to(torch.bfloat16)
    dqkv, cs = co.attention_bwd(dout, qkv, p, B, N, H, hd, want_colsum=True)
    dqkvr, csr = to.attention_bwd(dout.float(), qkv.float(), pr, B, N, H, hd, want_colsum=True)
    _close(dqkv, dqkvr)
    _close(cs, csr, rel=5e-2)
