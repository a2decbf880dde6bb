"""GPU: row f-4 -- single-vector scorer, the six bi-encoder losses (loss + every gradient) and similarity maps through the C
ABI (cpb_dense_dot_launch + cpb_colbert_loss_launch with d_q == NULL) against outputs of the reference
(tests/golden/bi_small.npz, oracle/make_golden.py::bi_small) and against the oracle port / fp32 torch at training shapes."""
import pytest
import torch

import colpali_b200 as cb
from conftest import load_golden
from oracle import li_oracle as O
from test_oracle_golden import BI_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

_CLASSES = {"ce": cb.BiEncoderLoss, "paired": cb.BiPairedEncoderLoss, "pairwise": cb.BiPairwiseCELoss,
            "sigmoid": cb.BiSigmoidLoss, "negce": cb.BiNegativeCELoss, "pairneg": cb.BiPairwiseNegativeCELoss}


def _run(mod, q, d, neg, off):
    qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
    nn = neg.clone().requires_grad_(True) if neg is not None else None
    loss = mod(qq, dd, nn, offset=off) if neg is not None else mod(qq, dd, offset=off)
    loss.backward()
    return loss.detach(), qq.grad, dd.grad, (nn.grad if nn is not None else None)


@pytest.mark.parametrize("case", BI_CASES, ids=[c[0] for c in BI_CASES])
def test_losses_against_reference_golden(case):
    name, sl, with_neg, off, kind, kw = case
    g = load_golden("bi_small.npz")
    q, d, neg = (torch.from_numpy(g[k]).to(DEV) for k in ("q", "d", "neg"))
    loss, dq, dd, dn = _run(_CLASSES[kind](**kw), q, d[sl].contiguous(), neg if with_neg else None, off)
    assert loss.dtype == torch.float32 and dq.dtype == q.dtype
    assert abs(float(loss) - float(g[f"{name}_loss"])) < 2e-5 * max(1.0, abs(float(g[f"{name}_loss"]))), name
    for got, key in ((dq, "dq"), (dd, "dd"), (dn, "dn")):
        if got is None:
            continue
        want = torch.from_numpy(g[f"{name}_{key}"])
        assert torch.allclose(got.cpu(), want, rtol=2e-4, atol=2e-5 * max(1.0, want.abs().max().item())), (name, key)


@pytest.mark.parametrize("kind,dtype,shape", [
    ("ce", torch.bfloat16, (64, 512, 1536, 448)),   # 8-rank gather of a BiQwen2 batch: C = 8 B, positives at rank 7
    ("paired", torch.float32, (48, 48, 1536, 0)),
    ("pairwise", torch.bfloat16, (64, 64, 2048, 0)),
    ("sigmoid", torch.float32, (32, 128, 1000, 64)),  # hidden size that is not a multiple of the 32-wide k stage
    ("negce", torch.bfloat16, (32, 64, 1536, 32)),
    ("pairneg", torch.float32, (24, 24, 777, 0)),
])
def test_losses_at_training_shapes_against_the_port(kind, dtype, shape):
    b, c, dim, off = shape
    g = torch.Generator().manual_seed(b + c)
    unit = lambda *sh: torch.nn.functional.normalize(torch.randn(*sh, generator=g), dim=-1)  # noqa: E731
    q, d = unit(b, dim), unit(c, dim)
    d[off : off + b] = torch.nn.functional.normalize(q + 3.0 * unit(b, dim), dim=-1)  # cos ~ 0.3: softmax not saturated
    neg = unit(b, 5, dim) if kind in ("negce", "pairneg") else None
    q, d = q.to(dtype), d.to(dtype)
    neg = neg.to(dtype) if neg is not None else None
    kw = dict(temperature=0.05, pos_aware_negative_filtering=(kind in ("ce", "sigmoid")))
    if kind == "pairneg":
        kw = dict(temperature=0.05)
    loss, dq, dd, dn = _run(_CLASSES[kind](**kw), q.to(DEV), d.to(DEV), neg.to(DEV) if neg is not None else None, off)
    # the port on fp32 copies of the same (possibly bf16) values: what the kernels compute -- exact operands, fp32 math
    qf, df = q.float().requires_grad_(True), d.float().requires_grad_(True)
    nf = neg.float().requires_grad_(True) if neg is not None else None
    want = O.bi_loss_port(kind, qf, df, nf, offset=off, **kw)
    want.backward()
    assert abs(float(loss) - float(want.detach())) < 1e-4 * max(1.0, abs(float(want.detach()))), (float(loss), float(want.detach()))
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-4  # gradients are returned in the embedding dtype
    for got, ref in ((dq, qf.grad), (dd, df.grad), (dn, nf.grad if nf is not None else None)):
        if got is None:
            continue
        assert got.dtype == dtype
        assert torch.allclose(got.float().cpu(), ref, rtol=tol, atol=tol * ref.abs().max().item())


def test_single_vector_scorer_keeps_the_operand_dtype():
    g = load_golden("bi_small.npz")
    q, d = torch.from_numpy(g["q"]), torch.from_numpy(g["d"])
    got = cb.score_single_vector(q, d, device=DEV)
    assert got.is_cuda and got.dtype == torch.float32
    assert torch.allclose(got.cpu(), torch.from_numpy(g["single_f32"]), rtol=1e-6, atol=1e-6)
    # bf16 lists: the reference rounds the result to bf16 before the fp32 cast, the kernel keeps the fp32 sum
    got16 = cb.score_single_vector(list(q.bfloat16()), list(d.bfloat16()), device=DEV)
    exact = torch.einsum("bd,cd->bc", q.bfloat16().float(), d.bfloat16().float())
    assert torch.allclose(got16.cpu(), exact, rtol=1e-6, atol=1e-6)
    assert torch.allclose(got16.cpu(), torch.from_numpy(g["single_bf16"]), atol=4e-3)  # bf16 spacing below 1
    with pytest.raises(ValueError, match="No queries"):
        cb.score_single_vector([], list(d), device=DEV)
    with pytest.raises(ValueError, match="No passages"):
        cb.score_single_vector(list(q), [], device=DEV)
    # a retrieval-sized call: 300 queries x 5000 pages x 1536 dims fp32 (the 64 x 64 tile variant)
    gen = torch.Generator(device=DEV).manual_seed(0)
    a = torch.randn(300, 1536, generator=gen, device=DEV)
    b = torch.randn(5000, 1536, generator=gen, device=DEV)
    want = a.double() @ b.double().t()
    assert torch.allclose(cb.score_single_vector(a, b, device=DEV).double(), want, rtol=1e-5, atol=1e-3)


def test_dense_dot_strides_gather_accumulate():
    gen = torch.Generator(device=DEV).manual_seed(1)
    a = torch.randn(37, 70, generator=gen, device=DEV)
    b = torch.randn(53, 70, generator=gen, device=DEV).bfloat16()
    want = a.double() @ b.double().t()
    assert torch.allclose(cb.dense_dot(a, b).double(), want, atol=1e-4)
    # transposed views (the backward products), no copies
    at, bt = a.t().contiguous().t(), b.t().contiguous().t()
    assert not at.is_contiguous() and torch.allclose(cb.dense_dot(at, bt).double(), want, atol=1e-4)
    # row gather + alpha + accumulate into a strided output
    rows = torch.tensor([5, 0, 52, 7, 7], dtype=torch.int32, device=DEV)
    out = torch.ones(37, 8, device=DEV)[:, :5]
    alpha = torch.tensor([0.5], device=DEV)
    cb.dense_dot(a, b, b_rows=rows, out=out, alpha=alpha, accumulate=True)
    assert torch.allclose(out.double(), 1 + 0.5 * want[:, rows.long()], atol=1e-4)


@pytest.mark.parametrize("m,n,k", [(37, 53, 72), (64, 512, 1536), (130, 300, 256), (8, 8, 2048), (256, 2304, 64), (40, 24, 1056)])
@pytest.mark.parametrize("dta,dtb", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                     (torch.float32, torch.bfloat16)], ids=["f32", "bf16", "f32xbf16"])
def test_dense_dot_tiled_paths(m, n, k, dta, dtb):
    """The vectorised kernels (64-tile with the K range split over a cluster, 128-tile) and the generic fallback, for
    k-contiguous operands and for transposed views (what the backward products pass), against fp64."""
    gen = torch.Generator(device=DEV).manual_seed(m * n + k)
    a = torch.randn(m, k, generator=gen, device=DEV).to(dta)
    b = torch.randn(n, k, generator=gen, device=DEV).to(dtb)
    want = a.double() @ b.double().t()
    tol = dict(rtol=1e-5, atol=2e-5 * k ** 0.5)
    at, bt = a.t().contiguous().t(), b.t().contiguous().t()   # same values, row-contiguous storage
    for x in (a, at):
        for y in (b, bt):
            assert torch.allclose(cb.dense_dot(x, y).double(), want, **tol), (x.stride(), y.stride())
    rows = torch.randint(0, n, (n + 3,), generator=gen, device=DEV).to(torch.int32)
    assert torch.allclose(cb.dense_dot(a, b, b_rows=rows).double(), want[:, rows.long()], **tol)
    # a retrieval-shaped launch takes the 128-tile kernel (more tiles than SMs)
    if (m, n, k) == (256, 2304, 64):
        big = torch.randn(1500, k, generator=gen, device=DEV).to(dta)
        huge = torch.randn(4000, k, generator=gen, device=DEV).to(dtb)
        assert torch.allclose(cb.dense_dot(big, huge).double(), big.double() @ huge.double().t(), **tol)


def test_similarity_maps_against_reference_golden():
    g = load_golden("bi_small.npz")
    img, qe, mask = (torch.from_numpy(g[k]).to(DEV) for k in ("map_img", "map_q", "map_mask"))
    maps = cb.get_similarity_maps_from_embeddings(img, qe, [(3, 4), (2, 5)], mask)
    for k, m in enumerate(maps):
        want = torch.from_numpy(g[f"map_{k}"])
        assert m.shape == want.shape and torch.allclose(m.cpu(), want, atol=1e-6)
    # broadcast n_patches tuple + the reference's sanity check (similarity_map_utils.py:34-40)
    same = cb.get_similarity_maps_from_embeddings(img[:1], qe[:1], (3, 4), mask[:1])
    assert torch.equal(same[0], maps[0])
    with pytest.raises(ValueError, match="does not match the number of non-padded image tokens"):
        cb.get_similarity_maps_from_embeddings(img, qe, (3, 4), mask)
    # ColQwen2-sized page: 32 x 32 patches of dim 128 in bf16 against 20 query tokens
    gen = torch.Generator(device=DEV).manual_seed(3)
    pe = torch.randn(1, 1030, 128, generator=gen, device=DEV).bfloat16()
    qq = torch.randn(1, 20, 128, generator=gen, device=DEV).bfloat16()
    pm = torch.zeros(1, 1030, dtype=torch.bool, device=DEV)
    pm[0, 4:1028] = True
    got = cb.get_similarity_maps_from_embeddings(pe, qq, (32, 32), pm)[0]
    want = O.similarity_maps_port(pe.float().cpu(), qq.float().cpu(), (32, 32), pm.cpu())[0]
    assert got.shape == (20, 32, 32) and torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-4)
