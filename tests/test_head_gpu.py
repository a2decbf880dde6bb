"""GPU parity: fused projection head against the reference model's own forward (tests/golden/head_small.npz,
captured from a random-init ColQwen2) and against the CPU oracle at production widths."""
import pytest
import torch

from conftest import from_bits, load_golden

import colpali_b200 as cb
from oracle import li_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bit_identical_fraction(a, b):
    return (a.view(torch.int16) == b.view(torch.int16)).float().mean().item()


def test_reference_model_forward_golden():
    g = load_golden("head_small.npz")
    mask = torch.from_numpy(g["bf16_mask"])
    h = from_bits(g["bf16_h"]).reshape(3, 24, -1)
    w = from_bits(g["bf16_w"]).reshape(128, -1)
    b = from_bits(g["bf16_b"])
    want = from_bits(g["bf16_out"]).reshape(3, 24, 128)
    got = cb.fused_head(h.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV)).cpu()
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    assert (got[mask == 0] == 0).all()                       # masked rows are zero rows
    assert bit_identical_fraction(got, want) > 0.995         # three-rounding emulation
    assert torch.allclose(got.float(), want.float(), rtol=0, atol=2 ** -8)  # never more than one bf16 ulp of |x| <= 1
    rows = got[mask == 1].float().norm(dim=-1)
    assert ((rows - 1).abs() < 8e-3).all()


@pytest.mark.parametrize("tokens,hidden", [
    (5000, 1536), (777, 2048), (256, 64),
    # shares of 64-token units per CTA: 6 or 7 units (three tiles + half a tile), 4 or 5, a single short tile, 64 + 1 rows
    (65920, 256), (38000, 512), (63, 64), (65, 128), (129, 64),
])
def test_production_widths_against_oracle(tokens, hidden):
    gen = torch.Generator().manual_seed(hidden)
    h = (torch.randn(tokens, hidden, generator=gen) * 2).bfloat16()
    w = (torch.randn(128, hidden, generator=gen) / hidden ** 0.5).bfloat16()
    b = (torch.randn(128, generator=gen) * 0.1).bfloat16()
    mask = (torch.rand(tokens, generator=gen) > 0.2).long()
    img = torch.rand(tokens, generator=gen) > 0.5
    want = O.head_port(h, w, b, mask)
    got = cb.fused_head(h.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV)).cpu()
    assert bit_identical_fraction(got, want) > 0.995
    assert torch.allclose(got.float(), want.float(), rtol=0, atol=2 ** -8)
    # image-token mask and the ModernVBert clamp variant
    want2 = O.head_port(h, w, b, mask, image_mask=img, clamp_norm=True)
    got2 = cb.fused_head(h.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV), img.to(DEV), clamp_norm=True).cpu()
    assert bit_identical_fraction(got2, want2) > 0.995
    assert (got2[(mask == 0) | ~img] == 0).all()
    # no bias, no mask, fp32-until-store variant is within half a bf16 ulp of the exact fp64 value
    exact = torch.nn.functional.normalize(h.double() @ w.double().T, dim=-1)
    got3 = cb.fused_head(h.to(DEV), w.to(DEV), None, single_rounding=True).cpu()
    assert torch.allclose(got3.double(), exact, rtol=0, atol=2 ** -9 + 1e-6)


def test_batched_shape_and_backward():
    gen = torch.Generator().manual_seed(1)
    h = (torch.randn(2, 40, 1536, generator=gen)).bfloat16().to(DEV).requires_grad_(True)
    lin = torch.nn.Linear(1536, 128).to(DEV, torch.bfloat16)
    mask = torch.ones(2, 40, dtype=torch.long, device=DEV)
    mask[0, :7] = 0
    out = cb.fused_head(h, lin.weight, lin.bias, mask)
    assert out.shape == (2, 40, 128)
    tgt = torch.randn(2, 40, 128, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    (out.float() * tgt).sum().backward()
    assert h.grad is not None and lin.weight.grad is not None and lin.bias.grad is not None
    # same gradients as the reference expression under autograd (bf16)
    h2 = h.detach().clone().requires_grad_(True)
    lin2 = torch.nn.Linear(1536, 128).to(DEV, torch.bfloat16)
    lin2.load_state_dict(lin.state_dict())
    ref = O.head_port(h2, lin2.weight, lin2.bias, mask)
    (ref.float() * tgt).sum().backward()
    assert torch.allclose(h.grad.float(), h2.grad.float(), rtol=5e-2, atol=1e-3)
    assert torch.allclose(lin.weight.grad.float(), lin2.weight.grad.float(), rtol=5e-2, atol=1e-2)
    assert (h.grad[0, :7] == 0).all()


def test_unsupported_shapes_fail_loudly():
    h = torch.zeros(4, 100, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(cb.ColpaliB200Error):
        cb.fused_head(h, torch.zeros(128, 100, dtype=torch.bfloat16, device=DEV), None)  # hidden % 64 != 0
    with pytest.raises(cb.ColpaliB200Error):
        cb.fused_head(torch.zeros(4, 128, dtype=torch.bfloat16, device=DEV),
                      torch.zeros(352, 128, dtype=torch.bfloat16, device=DEV), None)      # dim > 320
    with pytest.raises(cb.ColpaliB200Error):
        cb.fused_head(torch.zeros(4, 128, dtype=torch.bfloat16, device=DEV),
                      torch.zeros(200, 128, dtype=torch.bfloat16, device=DEV), None)      # dim % 32 != 0


@pytest.mark.parametrize("head_cluster", (1, 2), ids=("cluster1", "cluster2"))
@pytest.mark.parametrize("tokens,hidden,dim", [(5000, 2560, 320), (777, 2048, 320), (130, 64, 320), (1000, 1536, 192),
                                               (1000, 1536, 256), (127, 128, 160)])
def test_wide_projection_dims_against_oracle(tokens, hidden, dim, head_cluster):
    """head_wide_sm100.cu -- ColQwen3's dim = 320 (models/qwen3/colqwen3/modeling_colqwen3.py:48) and
    the other multiples of 32 above 128, with and without the 2-CTA W multicast."""
    from colpali_b200 import _lib
    gen = torch.Generator().manual_seed(hidden + dim)
    h = (torch.randn(tokens, hidden, generator=gen) * 2).bfloat16()
    w = (torch.randn(dim, hidden, generator=gen) / hidden ** 0.5).bfloat16()
    b = (torch.randn(dim, generator=gen) * 0.1).bfloat16()
    mask = (torch.rand(tokens, generator=gen) > 0.2).long()
    img = torch.rand(tokens, generator=gen) > 0.5
    _lib.set_option("head_cluster", head_cluster)
    try:
        want = O.head_port(h, w, b, mask)
        got = cb.fused_head(h.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV)).cpu()
        assert got.shape == (tokens, dim)
        assert bit_identical_fraction(got, want) > 0.995
        assert torch.allclose(got.float(), want.float(), rtol=0, atol=2 ** -8)
        want2 = O.head_port(h, w, b, mask, image_mask=img, clamp_norm=True)
        got2 = cb.fused_head(h.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV), img.to(DEV), clamp_norm=True).cpu()
        assert bit_identical_fraction(got2, want2) > 0.995
        assert (got2[(mask == 0) | ~img] == 0).all()
        exact = torch.nn.functional.normalize(h.double() @ w.double().T, dim=-1)
        got3 = cb.fused_head(h.to(DEV), w.to(DEV), None, single_rounding=True).cpu()
        assert torch.allclose(got3.double(), exact, rtol=0, atol=2 ** -9 + 1e-6)
    finally:
        _lib.set_option("head_cluster", 0)


def test_wide_dim320_reference_model_forward_golden():
    """head of a random-init reference ColQwen3 (dim 320), captured by oracle/make_golden.py wide."""
    g = load_golden("wide_dim320.npz")
    mask = torch.from_numpy(g["h_mask"])
    h = from_bits(g["h_h"]).reshape(3, 24, -1)
    w = from_bits(g["h_w"]).reshape(320, -1)
    b = from_bits(g["h_b"])
    want = from_bits(g["h_out"]).reshape(3, 24, 320)
    got = cb.fused_head(h.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV)).cpu()
    assert got.shape == want.shape and (got[mask == 0] == 0).all()
    assert bit_identical_fraction(got, want) > 0.995
    assert torch.allclose(got.float(), want.float(), rtol=0, atol=2 ** -8)
