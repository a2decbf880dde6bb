"""GPU parity: fused ColbertLoss / ColbertPairwiseCELoss (forward and backward) against the reference's own
outputs (tests/golden/loss_*.npz) and the CPU oracle."""
import math

import pytest
import torch

from conftest import load_golden

import colpali_b200 as cb
from oracle import li_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOSS_TOL = 1e-3  # BASELINE configs[2]: loss within 1e-3


def _run(mod, q, d, offset=0):
    qq = q.to(DEV).requires_grad_(True)
    dd = d.to(DEV).requires_grad_(True)
    loss = mod(qq, dd, offset=offset)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach().cpu(), qq.grad.cpu(), dd.grad.cpu()


CASES = (
    ("colbert", lambda: cb.ColbertLoss()),
    ("colbert_nonorm_t1", lambda: cb.ColbertLoss(temperature=1.0, normalize_scores=False)),
    ("colbert_filter", lambda: cb.ColbertLoss(pos_aware_negative_filtering=True)),
    ("pairwise", lambda: cb.ColbertPairwiseCELoss()),
    ("pairwise_filter", lambda: cb.ColbertPairwiseCELoss(pos_aware_negative_filtering=True)),
)


@pytest.mark.parametrize("name,make", CASES)
def test_small_losses_and_grads_match_reference(name, make):
    """B=4, C=6, N_q=5, N_d=9, D=16, offset=1, zero query rows and zero document rows; inputs are
    bf16-representable so the bf16 contraction is exact and the reference's fp32 numbers apply directly."""
    g = load_golden("loss_small.npz")
    q, d = torch.from_numpy(g["q"]), torch.from_numpy(g["d"])
    loss, dq, dd = _run(make(), q, d, offset=1)
    assert loss.dtype == torch.float32 and loss.dim() == 0
    assert abs(float(loss) - float(g[f"{name}_loss"])) < 2e-5, (float(loss), float(g[f"{name}_loss"]))
    # gradients: exclude all-zero (padding) rows, where amax's tie-splitting differs by construction
    real_q = q.abs().sum(-1) > 0
    real_d = d.abs().sum(-1) > 0
    ref_dq, ref_dd = torch.from_numpy(g[f"{name}_dq"]), torch.from_numpy(g[f"{name}_dd"])
    assert torch.allclose(dq[real_q], ref_dq[real_q], rtol=1e-4, atol=2e-6), name
    assert torch.allclose(dd[real_d], ref_dd[real_d], rtol=1e-4, atol=2e-6), name


def test_zero_embedding_known_answers():
    """The reference's own KATs (tests/loss/test_li_losses.py:76-100, :166-181): all-zero inputs."""
    b, nq, dim = 3, 1, 4
    q = torch.zeros(b, nq, dim, device=DEV)
    d = torch.zeros(b, nq, dim, device=DEV)
    ce = cb.ColbertLoss(temperature=1.0, normalize_scores=False)
    assert math.isclose(float(ce(q, d)), math.log(b), rel_tol=1e-6)
    filt = cb.ColbertLoss(temperature=1.0, normalize_scores=False, pos_aware_negative_filtering=True)
    assert math.isclose(float(filt(q, d)), float(ce(q, d)), rel_tol=1e-6)
    pw = cb.ColbertPairwiseCELoss(temperature=1.0, normalize_scores=False)
    assert math.isclose(float(pw(q, d)), math.log(2.0), rel_tol=1e-6)


def test_cfg3_loss_and_gradients():
    """B=64 pairs, N_q=32, documents 768..1030 tokens left-padded to 1030 (BASELINE configs[2])."""
    g = load_golden("loss_cfg3.npz")
    q, d, lens = O.cfg3_inputs()
    for name, mod in (("colbert", cb.ColbertLoss()), ("pairwise", cb.ColbertPairwiseCELoss())):
        loss, dq, dd = _run(mod, q, d)
        ref = float(g[f"{name}_fp32"])
        assert abs(float(loss) - ref) < LOSS_TOL, (name, float(loss), ref)
        ref_dq = torch.from_numpy(g[f"{name}_dq"])
        # grads come back in the embedding dtype (bf16): compare direction and scale
        cos = torch.nn.functional.cosine_similarity(dq.float().flatten(), ref_dq.flatten(), dim=0)
        assert cos > 0.9999, (name, float(cos))
        assert torch.allclose(dq.float(), ref_dq, rtol=2e-2, atol=ref_dq.abs().max().item() * 1e-2)
        ref_dd2 = torch.from_numpy(g[f"{name}_dd_first2"])
        real = d[:2].float().abs().sum(-1) > 0
        assert torch.allclose(dd[:2].float()[real], ref_dd2[real], rtol=2e-2, atol=ref_dd2.abs().max().item() * 1e-2)
        rn = dd.float().norm(dim=-1)
        ref_rn = torch.from_numpy(g[f"{name}_dd_rownorm"])
        realall = d.float().abs().sum(-1) > 0
        assert torch.allclose(rn[realall], ref_rn[realall], rtol=3e-2, atol=ref_rn.max().item() * 1e-2)


def test_gathered_documents_offset_and_no_grad_path():
    """C > B with a rank offset (contrastive_trainer.py:143-150 gathers documents across ranks)."""
    g = torch.Generator().manual_seed(4)
    q = O.unit_rows((8, 20, 128), 10)
    d = O.unit_rows((24, 70, 128), 11)
    d[:, :5] = 0
    for offset in (0, 8, 16):
        want = O.colbert_loss_port(q.float(), d.float(), offset=offset)
        got = cb.ColbertLoss()(q.to(DEV), d.to(DEV), offset=offset)  # no grad required: argmax-free kernel
        assert abs(float(got) - float(want)) < 1e-4
        want_p = O.colbert_pairwise_ce_loss_port(q.float(), d.float(), offset=offset)
        got_p = cb.ColbertPairwiseCELoss()(q.to(DEV), d.to(DEV), offset=offset)
        assert abs(float(got_p) - float(want_p)) < 1e-4
    with pytest.raises(cb.ColpaliB200Error):
        cb.ColbertLoss()(q.to(DEV), d.to(DEV), offset=17)  # positive index past the last document


def test_upstream_gradient_scaling_and_autograd_of_oracle():
    """d(2.5 * loss) = 2.5 * d(loss); gradients agree with torch autograd through the oracle port."""
    q = O.unit_rows((6, 32, 128), 20)
    d = O.unit_rows((6, 300, 128), 21)
    qq, dd = q.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    (cb.ColbertLoss()(qq, dd) * 2.5).backward()
    qo, do = q.float().requires_grad_(True), d.float().requires_grad_(True)
    (O.colbert_loss_port(qo, do) * 2.5).backward()
    assert torch.allclose(qq.grad.float().cpu(), qo.grad, rtol=2e-2, atol=qo.grad.abs().max().item() * 1e-2)
    assert torch.allclose(dd.grad.float().cpu(), do.grad, rtol=2e-2, atol=do.grad.abs().max().item() * 1e-2)


SMOOTH_CASES = (
    ("colbert", lambda: cb.ColbertLoss(use_smooth_max=True)),
    ("colbert_tau05_nonorm", lambda: cb.ColbertLoss(use_smooth_max=True, tau=0.05, normalize_scores=False, temperature=0.5)),
    ("pairwise", lambda: cb.ColbertPairwiseCELoss(use_smooth_max=True)),
    ("pairwise_filter", lambda: cb.ColbertPairwiseCELoss(use_smooth_max=True, pos_aware_negative_filtering=True)),
)


def _close_grad(got, want, what):
    # the softmax weights of the smooth-max backward are rounded to bf16 before the second tensor-core product
    assert torch.allclose(got, want, rtol=2e-2, atol=want.abs().max().item() * 1e-2), what


@pytest.mark.parametrize("name,make", SMOOTH_CASES)
def test_smooth_max_small_losses_and_grads_match_reference(name, make):
    """use_smooth_max=True (tau * logsumexp over document tokens, late_interaction_losses.py:40-44, :88-90) against the
    reference's loss, dq and dd.  Zero query rows count (each adds tau * log N_d) and zero document rows take part in
    the log-sum-exp, so -- unlike the hard max -- every gradient row is comparable."""
    g = load_golden("loss_smooth.npz")
    q, d = torch.from_numpy(g["q"]), torch.from_numpy(g["d"])
    loss, dq, dd = _run(make(), q, d, offset=1)
    assert abs(float(loss) - float(g[f"{name}_loss"])) < 1e-4, (float(loss), float(g[f"{name}_loss"]))
    _close_grad(dq, torch.from_numpy(g[f"{name}_dq"]), (name, "dq"))
    _close_grad(dd, torch.from_numpy(g[f"{name}_dd"]), (name, "dd"))


def test_smooth_max_sigmoid_and_negative_losses_match_reference():
    g = load_golden("loss_smooth.npz")
    q, d, neg = (torch.from_numpy(g[k]) for k in ("q", "d", "neg"))
    loss, dq, dd = _run(cb.ColbertSigmoidLoss(use_smooth_max=True), q, d[:4])
    assert abs(float(loss) - float(g["sigmoid_loss"])) < 1e-4
    _close_grad(dq, torch.from_numpy(g["sigmoid_dq"]), "sigmoid dq")
    _close_grad(dd, torch.from_numpy(g["sigmoid_dd"]), "sigmoid dd")
    for name, mod in (("negce", cb.ColbertNegativeCELoss(use_smooth_max=True)),
                      ("pairneg", cb.ColbertPairwiseNegativeCELoss(use_smooth_max=True, in_batch_term_weight=0.3))):
        qq, dd_, nn = (t.to(DEV).requires_grad_(True) for t in (q, d, neg))
        loss = mod(qq, dd_, nn, offset=1)
        loss.backward()
        assert abs(loss.item() - float(g[f"{name}_loss"])) < 1e-4, name
        for got, key in ((qq.grad, "dq"), (dd_.grad, "dd"), (nn.grad, "dn")):
            _close_grad(got.cpu(), torch.from_numpy(g[f"{name}_{key}"]), (name, key))


def test_smooth_max_cfg3_loss_and_gradients():
    """BASELINE configs[2] shapes with use_smooth_max=True: loss within 1e-3 of the fp32 reference, gradients by
    direction and scale (they come back in bf16)."""
    g = load_golden("loss_smooth.npz")
    q, d, _ = O.cfg3_inputs()
    assert abs(float(q.double().sum()) - float(g["cfg3_q_checksum"])) < 1e-6
    for name, mod in (("colbert", cb.ColbertLoss(use_smooth_max=True)), ("pairwise", cb.ColbertPairwiseCELoss(use_smooth_max=True))):
        loss, dq, dd = _run(mod, q, d)
        ref = float(g[f"cfg3_{name}_fp32"])
        assert abs(float(loss) - ref) < LOSS_TOL, (name, float(loss), ref)
        ref_dq = torch.from_numpy(g[f"cfg3_{name}_dq"])
        cos = torch.nn.functional.cosine_similarity(dq.float().flatten(), ref_dq.flatten(), dim=0)
        assert cos > 0.999, (name, float(cos))
        assert torch.allclose(dq.float(), ref_dq, rtol=3e-2, atol=ref_dq.abs().max().item() * 2e-2)
        ref_dd2 = torch.from_numpy(g[f"cfg3_{name}_dd_first2"])
        cos = torch.nn.functional.cosine_similarity(dd[:2].float().flatten(), ref_dd2.flatten(), dim=0)
        assert cos > 0.999, (name, float(cos))
        rn, ref_rn = dd.float().norm(dim=-1), torch.from_numpy(g[f"cfg3_{name}_dd_rownorm"])
        assert torch.allclose(rn, ref_rn, rtol=5e-2, atol=ref_rn.max().item() * 2e-2)


def test_smooth_max_unsupported_combinations_are_refused_loudly():
    q = O.unit_rows((2, 4, 320), 1).to(DEV).requires_grad_(True)
    with pytest.raises(cb.ColpaliB200Error):  # the smooth-max backward serves embedding dim 128 only
        cb.ColbertLoss(use_smooth_max=True)(q, q).backward()


def test_long_queries_take_the_two_kernel_path():
    """N_q = 40 -> nq_pad = 64: segment sums + stand-alone loss kernel; same numbers as the oracle."""
    q = O.unit_rows((5, 40, 128), 40)
    d = O.unit_rows((5, 200, 128), 41)
    loss, dq, dd = _run(cb.ColbertLoss(), q, d)
    qo, do = q.float().requires_grad_(True), d.float().requires_grad_(True)
    ref = O.colbert_loss_port(qo, do)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4
    assert torch.allclose(dq.float(), qo.grad, rtol=2e-2, atol=qo.grad.abs().max().item() * 1e-2)
    assert torch.allclose(dd.float(), do.grad, rtol=2e-2, atol=do.grad.abs().max().item() * 1e-2)


NEG_CASES = (
    ("negce", lambda: cb.ColbertNegativeCELoss()),
    ("negce_w0", lambda: cb.ColbertNegativeCELoss(in_batch_term_weight=0.0)),
    ("negce_filter", lambda: cb.ColbertNegativeCELoss(pos_aware_negative_filtering=True, in_batch_term_weight=0.3)),
    ("pairneg", lambda: cb.ColbertPairwiseNegativeCELoss()),
    ("pairneg_t1", lambda: cb.ColbertPairwiseNegativeCELoss(temperature=1.0, in_batch_term_weight=0.7)),
)


@pytest.mark.parametrize("name,make", NEG_CASES)
def test_explicit_negative_losses_match_reference(name, make):
    """ColbertNegativeCELoss / ColbertPairwiseNegativeCELoss (late_interaction_losses.py:215-252, :361-398):
    B=4, C=6 (offset 1), 3 negatives per query, zero rows; losses and all three gradients vs the reference."""
    g = load_golden("loss_neg_small.npz")
    q, d, neg = (torch.from_numpy(g[k]) for k in ("q", "d", "neg"))
    qq, dd, nn = (t.to(DEV).requires_grad_(True) for t in (q, d, neg))
    loss = make()(qq, dd, nn, offset=1)
    loss.backward()
    assert abs(loss.item() - float(g[f"{name}_loss"])) < 2e-5, (loss.item(), float(g[f"{name}_loss"]))
    for got, key, src in ((qq.grad, "dq", q), (dd.grad, "dd", d), (nn.grad, "dn", neg)):
        real = src.abs().sum(-1) > 0
        ref = torch.from_numpy(g[f"{name}_{key}"])
        assert torch.allclose(got.cpu()[real], ref[real], rtol=1e-4, atol=2e-6), (name, key)


@pytest.mark.parametrize("name,make", (("sigmoid", lambda: cb.ColbertSigmoidLoss()),
                                       ("sigmoid_filter_t1", lambda: cb.ColbertSigmoidLoss(temperature=1.0, pos_aware_negative_filtering=True))))
def test_sigmoid_loss_matches_reference(name, make):
    g = load_golden("loss_neg_small.npz")
    q, d = torch.from_numpy(g["q"]), torch.from_numpy(g["d"])[:4]
    loss, dq, dd = _run(make(), q, d)
    assert abs(float(loss) - float(g[f"{name}_loss"])) < 2e-5
    real_q, real_d = q.abs().sum(-1) > 0, d.abs().sum(-1) > 0
    assert torch.allclose(dq[real_q], torch.from_numpy(g[f"{name}_dq"])[real_q], rtol=1e-4, atol=2e-6)
    assert torch.allclose(dd[real_d], torch.from_numpy(g[f"{name}_dd"])[real_d], rtol=1e-4, atol=2e-6)
    with pytest.raises(ValueError):
        cb.ColbertSigmoidLoss()(q.to(DEV), torch.from_numpy(g["d"]).to(DEV))  # non-square score matrix


@pytest.mark.parametrize("kind", ("negce", "negce_filter", "pairneg", "sigmoid"))
def test_explicit_negative_and_sigmoid_losses_at_training_shapes(kind):
    """the a5 losses were only checked at B = 4.  B = 32 queries of 32 tokens, C = 64 gathered
    documents (offset 32) of 200..300 tokens left-padded with zero rows, 3 negatives per query, bf16 -- against the
    pinned oracle port evaluated in fp32 on copies of the same bf16 values (losses) and through autograd (gradients)."""
    gen = torch.Generator().manual_seed(31)
    b, c, n_neg, nq, nd, off = 32, 64, 3, 32, 300, 32
    unit = lambda *sh: torch.nn.functional.normalize(torch.randn(*sh, generator=gen), dim=-1)  # noqa: E731
    q, d, neg = unit(b, nq, 128), unit(c, nd, 128), unit(b, n_neg, nd, 128)
    q[:, 24:] = 0                                                   # 24 real query tokens (the rest is padding)
    d[off:off + b, -nq:] = torch.nn.functional.normalize(q + 0.5 * unit(b, nq, 128), dim=-1)  # planted positives
    d[off:off + b, -nq:][:, 24:] = unit(b, 8, 128)
    lens = torch.randint(200, nd + 1, (c,), generator=gen)
    for j in range(c):
        d[j, : nd - int(lens[j])] = 0
    neg[:, :, :40] = 0
    q, d, neg = q.bfloat16(), d.bfloat16(), neg.bfloat16()
    if kind == "sigmoid":
        d, off = d[off:off + b].contiguous(), 0
        mod, port, kw = cb.ColbertSigmoidLoss(), O.colbert_sigmoid_loss_port, {}
    elif kind == "pairneg":  # (temperatures at which the planted positives do not saturate the softplus / softmax:
        # at the default 0.02 every gradient of this batch is ~1e-12 and a comparison of directions is noise)
        kw = dict(temperature=0.5, in_batch_term_weight=0.4)
        mod, port = cb.ColbertPairwiseNegativeCELoss(**kw), O.colbert_pairwise_negative_ce_loss_port
    else:
        kw = dict(temperature=0.3)
        if kind == "negce_filter":
            kw.update(pos_aware_negative_filtering=True, in_batch_term_weight=0.3)
        mod, port = cb.ColbertNegativeCELoss(**kw), O.colbert_negative_ce_loss_port
    qq, dd = q.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    qo, do = q.float().requires_grad_(True), d.float().requires_grad_(True)
    if kind == "sigmoid":
        loss, want = mod(qq, dd), port(qo, do)
        nn = no = None
    else:
        nn, no = neg.to(DEV).requires_grad_(True), neg.float().requires_grad_(True)
        loss, want = mod(qq, dd, nn, offset=off), port(qo, do, no, offset=off, **kw)
    loss.backward()
    want.backward()
    assert abs(float(loss) - float(want.detach())) < LOSS_TOL, (kind, float(loss), float(want.detach()))
    for got, ref, src in ((qq.grad, qo.grad, q), (dd.grad, do.grad, d), (nn.grad if nn is not None else None, no.grad if no is not None else None, neg)):
        if got is None:
            continue
        real = (src.float().abs().sum(-1) > 0)  # zero (padding) rows tie everywhere: amax splits, the kernel picks one (a8)
        g_, r_ = got.float().cpu()[real], ref[real]
        cos = torch.nn.functional.cosine_similarity(g_.flatten(), r_.flatten(), dim=0)
        assert cos > 0.999, (kind, float(cos))
        assert torch.allclose(g_, r_, rtol=3e-2, atol=r_.abs().max().item() * 2e-2), kind


def test_reference_kats_for_negative_losses():
    """tests/loss/test_li_losses.py:103-181: all-zero embeddings -> softplus(0) = ln 2 (with and without in-batch term)."""
    b, nq, dim, nneg = 2, 1, 3, 1
    q = torch.zeros(b, nq, dim, device=DEV)
    d = torch.zeros(b, nq, dim, device=DEV)
    n = torch.zeros(b, nneg, nq, dim, device=DEV)
    ln2 = math.log(2.0)
    no_ib = cb.ColbertNegativeCELoss(temperature=1.0, normalize_scores=False, in_batch_term_weight=0)
    assert math.isclose(float(no_ib(q, d, n)), ln2, rel_tol=1e-6)
    with_ib = cb.ColbertNegativeCELoss(temperature=1.0, normalize_scores=False, in_batch_term_weight=0.5)
    assert math.isclose(float(with_ib(q, d, n)), ln2, rel_tol=1e-6)  # in-batch CE over 2 zeros is ln 2 as well
    pw = cb.ColbertPairwiseNegativeCELoss(temperature=1.0, normalize_scores=False, in_batch_term_weight=0.5)
    assert math.isclose(float(pw(q, d, n)), ln2, rel_tol=1e-6)


@pytest.mark.parametrize("dim", (192, 256, 320))
def test_wide_embeddings_loss_and_gradients(dim):
    """ColQwen3-style embedding dims through the K-pipelined scorer and the dim-generic backward kernels;
    loss and gradients agree with torch autograd through the oracle port."""
    q = O.unit_rows((6, 20, dim), 30 + dim)
    d = O.unit_rows((6, 300, dim), 31 + dim)
    neg = O.unit_rows((6, 2, 150, dim), 32 + dim)
    qq, dd = q.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    loss = cb.ColbertLoss()(qq, dd)
    loss.backward()
    qo, do = q.float().requires_grad_(True), d.float().requires_grad_(True)
    ref = O.colbert_loss_port(qo, do)
    ref.backward()
    assert abs(float(loss) - float(ref)) < LOSS_TOL
    assert qq.grad.shape == q.shape and dd.grad.shape == d.shape
    assert torch.allclose(qq.grad.float().cpu(), qo.grad, rtol=2e-2, atol=qo.grad.abs().max().item() * 1e-2)
    assert torch.allclose(dd.grad.float().cpu(), do.grad, rtol=2e-2, atol=do.grad.abs().max().item() * 1e-2)
    # explicit negatives: both score matrices and all three gradients
    qq, dd, nn = (t.to(DEV).requires_grad_(True) for t in (q, d, neg))
    loss = cb.ColbertNegativeCELoss()(qq, dd, nn)
    loss.backward()
    qo, do, no = (t.float().requires_grad_(True) for t in (q, d, neg))
    ref = O.colbert_negative_ce_loss_port(qo, do, no)
    ref.backward()
    assert abs(float(loss) - float(ref)) < LOSS_TOL
    for got, want in ((qq.grad, qo.grad), (dd.grad, do.grad), (nn.grad, no.grad)):
        assert torch.allclose(got.float().cpu(), want, rtol=2e-2, atol=want.abs().max().item() * 1e-2)


@pytest.mark.parametrize("name,make", (("colbert", lambda: cb.ColbertLoss()), ("pairwise", lambda: cb.ColbertPairwiseCELoss())))
def test_wide_dim320_losses_against_reference_golden(name, make):
    """Loss and gradients at dim 320 against the reference's own numbers (bf16-representable inputs)."""
    g = load_golden("wide_dim320.npz")
    q, d = torch.from_numpy(g["l_q"]), torch.from_numpy(g["l_d"])
    loss, dq, dd = _run(make(), q, d, offset=1)
    assert abs(float(loss) - float(g[f"l_{name}_loss"])) < 2e-5
    real_q, real_d = q.abs().sum(-1) > 0, d.abs().sum(-1) > 0
    assert torch.allclose(dq[real_q], torch.from_numpy(g[f"l_{name}_dq"])[real_q], rtol=1e-4, atol=2e-6)
    assert torch.allclose(dd[real_d], torch.from_numpy(g[f"l_{name}_dd"])[real_d], rtol=1e-4, atol=2e-6)


def test_bf16_gradient_rows_are_the_cast_of_the_fp32_rows():
    """CPB_FLAG_GRAD_BF16: the backward kernels round each gradient row themselves; bit-identical to casting the fp32 rows."""
    from colpali_b200 import losses as L

    torch.manual_seed(5)
    q = torch.randn(6, 20, 128).bfloat16()
    d = torch.randn(6, 300, 128).bfloat16()
    d[1, :40] = 0
    grads = {}
    orig = L._maxsim_backward
    try:
        for bf16 in (False, True):
            L._maxsim_backward = lambda *a, _o=orig, _b=bf16, **k: _o(*a, **{**k, "bf16_out": _b})
            _, dq, dd = _run(cb.ColbertLoss(), q, d)
            grads[bf16] = (dq, dd)
    finally:
        L._maxsim_backward = orig
    assert grads[True][0].dtype == torch.bfloat16
    assert torch.equal(grads[True][0], grads[False][0]) and torch.equal(grads[True][1], grads[False][1])


def test_fused_loss_workspace_survives_changing_batch_sizes():
    """The per-stream counter workspace of the fused loss is shared by launches of different shapes: 1, 5, 2 and 8
    query-tile groups in turn (a first version packed the groups' partial sums behind the counters, so a launch with more
    groups than its predecessor started from non-zero counters and never emitted its loss)."""
    gen = torch.Generator().manual_seed(77)
    unit = lambda *sh: torch.nn.functional.normalize(torch.randn(*sh, generator=gen), dim=-1).bfloat16()  # noqa: E731
    for rep in range(2):
        for b in (4, 40, 12, 64, 3):
            q, d = unit(b, 32, 128), unit(b + 5, 90, 128)
            for mod, port in ((cb.ColbertLoss(), O.colbert_loss_port), (cb.ColbertPairwiseCELoss(), O.colbert_pairwise_ce_loss_port)):
                got = mod(q.to(DEV).requires_grad_(True), d.to(DEV))      # argmax kernel + per-group loss tail
                got_ng = mod(q.to(DEV), d.to(DEV))                         # max kernel, no gradient
                want = port(q.float(), d.float())
                assert abs(float(got) - float(want)) < 2e-4, (rep, b, float(got), float(want))
                assert abs(float(got_ng) - float(want)) < 2e-4, (rep, b, float(got_ng), float(want))


def test_back_to_back_training_forwards_do_not_deadlock():
    """400 argmax-mode forwards without host synchronisation (a training loop): the epilogue warpgroups of every CTA must
    get their registers (setmaxnreg) in every launch -- a race there killed one launch in a few hundred."""
    q, d, _ = O.cfg3_inputs()
    q, d = q.to(DEV).requires_grad_(True), d.to(DEV).requires_grad_(True)
    mod = cb.ColbertLoss()
    losses = [mod(q, d).detach() for _ in range(400)]
    torch.cuda.synchronize()
    assert all(torch.equal(x, losses[0]) for x in losses)
    q.grad = None
    d.grad = None
    for _ in range(100):
        mod(q, d).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(q.grad).all() and torch.isfinite(d.grad).all()
