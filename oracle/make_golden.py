"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Every fixture stores the reference's outputs; inputs are either stored (small cases) or re-created
from the seeded generators in oracle/li_oracle.py with a checksum to detect RNG drift.  The same
script asserts that the oracle port agrees with the reference on every case, which is what pins the
oracle ("parity pinned against reference outputs").
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from colpali_engine.loss import (  # noqa: E402  (reference)
    ColbertLoss,
    ColbertNegativeCELoss,
    ColbertPairwiseCELoss,
    ColbertPairwiseNegativeCELoss,
)
from colpali_engine.loss.late_interaction_losses import ColbertSigmoidLoss  # noqa: E402
from colpali_engine.utils.processing_utils import BaseVisualRetrieverProcessor as RefProc  # noqa: E402

from oracle import li_oracle as O  # noqa: E402

# CPB_GOLDEN_OUT=/tmp/somewhere regenerates into another directory (oracle/compare_golden.py then diffs it against the
# committed fixtures: how reproducible the reference's CPU outputs are on THIS host)
GOLD = os.environ.get("CPB_GOLDEN_OUT") or os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def bits(x: torch.Tensor) -> np.ndarray:
    """bf16 tensor -> uint16 bit patterns (npz has no bf16)."""
    return x.contiguous().view(torch.int16).numpy().view(np.uint16)


def checksum(x: torch.Tensor) -> float:
    return float(x.double().sum())


def ref_score(qs, ps, **kw):
    return RefProc.score_multi_vector(qs, ps, device="cpu", **kw)


def scorer_small():
    out = {}
    # (1) the reference's own unit test shapes, tests/utils/test_processing_utils.py:15-35
    g = torch.Generator().manual_seed(7)
    qs = [torch.randn(2, 32, generator=g), torch.randn(4, 32, generator=g)]
    ps = [torch.randn(8, 32, generator=g), torch.randn(4, 32, generator=g), torch.randn(16, 32, generator=g)]
    out["t1_q"] = torch.nn.utils.rnn.pad_sequence(qs, batch_first=True).numpy()
    out["t1_qlen"] = np.array([2, 4])
    out["t1_p"] = torch.nn.utils.rnn.pad_sequence(ps, batch_first=True).numpy()
    out["t1_plen"] = np.array([8, 4, 16])
    out["t1_list"] = ref_score(qs, ps).numpy()
    out["t1_tensor"] = ref_score(torch.nn.utils.rnn.pad_sequence(qs, batch_first=True),
                                 torch.nn.utils.rnn.pad_sequence(ps, batch_first=True)).numpy()
    assert np.allclose(out["t1_list"], out["t1_tensor"])
    assert torch.allclose(O.score_multi_vector_port(qs, ps), torch.from_numpy(out["t1_list"]))

    # (2) the zero-padding trap of SURVEY.md section 8 a1' (4)
    q = [torch.tensor([[1.0, 0.0]])]
    a = torch.tensor([[-1.0, 0.0], [-0.5, 0.0]])
    b3 = torch.tensor([[-0.2, 0.0], [-0.3, 0.0], [-0.9, 0.0]])
    out["t2_alone"] = ref_score(q, [a]).numpy()
    out["t2_batched"] = ref_score(q, [a, b3]).numpy()
    out["t2_bs1"] = ref_score(q, [a, b3], batch_size=1).numpy()
    assert out["t2_alone"][0, 0] == -0.5 and out["t2_batched"][0, 0] == 0.0 and out["t2_bs1"][0, 0] == -0.5

    # (3) ragged bf16, several padding groups (batch_size=3) and the default grouping
    g = torch.Generator().manual_seed(11)
    qlen = [5, 32, 17, 1, 40]
    plen = [1, 16, 17, 255, 256, 257, 600, 1030, 7, 64]
    qs = [torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).bfloat16() for n in qlen]
    ps = [torch.nn.functional.normalize(torch.randn(n, 128, generator=g), dim=-1).bfloat16() for n in plen]
    out["t3_q"] = bits(torch.cat(qs))
    out["t3_qlen"] = np.array(qlen)
    out["t3_p"] = bits(torch.cat(ps))
    out["t3_plen"] = np.array(plen)
    out["t3_bf16"] = ref_score(qs, ps).numpy()
    out["t3_fp32"] = ref_score([x.float() for x in qs], [x.float() for x in ps]).numpy()
    out["t3_fp32_bs3"] = ref_score([x.float() for x in qs], [x.float() for x in ps], batch_size=3).numpy()
    for bs, key in ((128, "t3_fp32"), (3, "t3_fp32_bs3")):
        port = O.score_multi_vector_port([x.float() for x in qs], [x.float() for x in ps], batch_size=bs)
        assert torch.equal(port, torch.from_numpy(out[key])), key
        f64 = O.maxsim_f64(qs, ps, O.reference_floors(plen, bs))
        assert np.allclose(f64, out[key], rtol=1e-5, atol=1e-5), key
    assert torch.equal(O.score_multi_vector_port(qs, ps), torch.from_numpy(out["t3_bf16"]))
    np.savez_compressed(os.path.join(GOLD, "scorer_small.npz"), **out)
    print("scorer_small ok")


def scorer_cfg(name, q, d):
    t = time.time()
    bf = ref_score(q, d)
    t_bf = time.time() - t
    t = time.time()
    fp = ref_score(q.float(), d.float())
    t_fp = time.time() - t
    assert torch.equal(O.score_multi_vector_port(q, d), bf)
    port_fp = O.score_multi_vector_port(q.float(), d.float())
    assert torch.equal(port_fp, fp)
    sub = slice(0, 16)
    f64 = O.maxsim_f64(list(q), list(d[sub]))
    assert np.allclose(f64, fp[:, sub].numpy(), rtol=2e-6, atol=2e-5)
    np.savez_compressed(
        os.path.join(GOLD, f"scorer_{name}.npz"),
        ref_bf16=bf.numpy(), ref_fp32=fp.numpy(),
        q_checksum=checksum(q), d_checksum=checksum(d),
        q_shape=np.array(q.shape), d_shape=np.array(d.shape),
        ref_seconds=np.array([t_bf, t_fp]), threads=torch.get_num_threads(),
    )
    print(f"scorer_{name} ok (reference CPU {t_bf:.2f}s bf16, {t_fp:.2f}s fp32)")


def loss_small():
    """B=4, C=6, N_q=5, N_d=9, fp32, offset=1, zero query rows and zero doc rows."""
    g = torch.Generator().manual_seed(5)
    # values are rounded to bf16 so that a bf16-contracting implementation sees exactly the same numbers
    q = torch.nn.functional.normalize(torch.randn(4, 5, 16, generator=g), dim=-1).bfloat16().float()
    d = torch.nn.functional.normalize(torch.randn(6, 9, 16, generator=g), dim=-1).bfloat16().float()
    q[1, 3:] = 0
    q[3, 4:] = 0
    d[0, :2] = 0
    d[4, :1] = 0
    out = {"q": q.numpy().copy(), "d": d.numpy().copy()}
    for name, mod, kw in (
        ("colbert", ColbertLoss(), {}),
        ("colbert_nonorm_t1", ColbertLoss(temperature=1.0, normalize_scores=False), {}),
        ("colbert_filter", ColbertLoss(pos_aware_negative_filtering=True), {}),
        ("pairwise", ColbertPairwiseCELoss(), {}),
        ("pairwise_filter", ColbertPairwiseCELoss(pos_aware_negative_filtering=True), {}),
    ):
        qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
        loss = mod(qq, dd, offset=1)
        loss.backward()
        out[f"{name}_loss"] = loss.detach().numpy()
        out[f"{name}_dq"] = qq.grad.numpy()
        out[f"{name}_dd"] = dd.grad.numpy()
    # port agreement
    assert torch.allclose(O.colbert_loss_port(q, d, offset=1), torch.from_numpy(out["colbert_loss"]))
    assert torch.allclose(O.colbert_loss_port(q, d, offset=1, pos_aware_negative_filtering=True),
                          torch.from_numpy(out["colbert_filter_loss"]))
    assert torch.allclose(O.colbert_pairwise_ce_loss_port(q, d, offset=1), torch.from_numpy(out["pairwise_loss"]))
    np.savez_compressed(os.path.join(GOLD, "loss_small.npz"), **out)
    print("loss_small ok")


def loss_neg_small():
    """Explicit-negative and sigmoid losses: B=4, C=6 (offset 1), n_neg=3, N_q=5, N_d=9 / 7, bf16-representable fp32."""
    g = torch.Generator().manual_seed(6)
    rnd = lambda *sh: torch.nn.functional.normalize(torch.randn(*sh, generator=g), dim=-1).bfloat16().float()  # noqa: E731
    q, d, neg = rnd(4, 5, 16), rnd(6, 9, 16), rnd(4, 3, 7, 16)
    q[2, 4:] = 0
    d[1, :3] = 0
    neg[0, 1, :2] = 0
    out = {"q": q.numpy().copy(), "d": d.numpy().copy(), "neg": neg.numpy().copy()}
    cases = (
        ("negce", ColbertNegativeCELoss(), O.colbert_negative_ce_loss_port, {}),
        ("negce_w0", ColbertNegativeCELoss(in_batch_term_weight=0.0), O.colbert_negative_ce_loss_port, dict(in_batch_term_weight=0.0)),
        ("negce_filter", ColbertNegativeCELoss(pos_aware_negative_filtering=True, in_batch_term_weight=0.3),
         O.colbert_negative_ce_loss_port, dict(pos_aware_negative_filtering=True, in_batch_term_weight=0.3)),
        ("pairneg", ColbertPairwiseNegativeCELoss(), O.colbert_pairwise_negative_ce_loss_port, {}),
        ("pairneg_t1", ColbertPairwiseNegativeCELoss(temperature=1.0, in_batch_term_weight=0.7),
         O.colbert_pairwise_negative_ce_loss_port, dict(temperature=1.0, in_batch_term_weight=0.7)),
    )
    for name, mod, port, kw in cases:
        qq, dd, nn = q.clone().requires_grad_(True), d.clone().requires_grad_(True), neg.clone().requires_grad_(True)
        loss = mod(qq, dd, nn, offset=1)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_dq"], out[f"{name}_dd"], out[f"{name}_dn"] = (
            loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy(), nn.grad.numpy())
        assert torch.allclose(port(q, d, neg, offset=1, **kw), loss.detach(), atol=1e-6), name
    for name, mod, kw in (("sigmoid", ColbertSigmoidLoss(), {}),
                          ("sigmoid_filter_t1", ColbertSigmoidLoss(temperature=1.0, pos_aware_negative_filtering=True),
                           dict(temperature=1.0, pos_aware_negative_filtering=True))):
        qq, dd = q.clone().requires_grad_(True), d[:4].clone().requires_grad_(True)
        loss = mod(qq, dd)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_dq"], out[f"{name}_dd"] = loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy()
        assert torch.allclose(O.colbert_sigmoid_loss_port(q, d[:4], **kw), loss.detach(), atol=1e-6), name
    np.savez_compressed(os.path.join(GOLD, "loss_neg_small.npz"), **out)
    print("loss_neg_small ok")


def loss_cfg3():
    q, d, lens = O.cfg3_inputs()
    out = {"q_checksum": checksum(q), "d_checksum": checksum(d), "lens": lens.numpy()}
    for name, cls in (("colbert", ColbertLoss), ("pairwise", ColbertPairwiseCELoss)):
        out[f"{name}_bf16"] = cls()(q, d).float().numpy()
        qq, dd = q.float().requires_grad_(True), d.float().requires_grad_(True)
        t = time.time()
        loss = cls()(qq, dd)
        out[f"{name}_fwd_seconds"] = time.time() - t
        loss.backward()
        out[f"{name}_fp32"] = loss.detach().numpy()
        out[f"{name}_dq"] = qq.grad.numpy()
        out[f"{name}_dd_first2"] = dd.grad[:2].numpy()
        out[f"{name}_dd_abs_sum"] = float(dd.grad.abs().sum())
        out[f"{name}_dd_rownorm"] = dd.grad.norm(dim=-1).numpy().astype(np.float32)
        print(f"  {name}: bf16 {float(out[name + '_bf16']):.5f} fp32 {float(out[name + '_fp32']):.5f}")
    assert torch.allclose(O.colbert_loss_port(q.float(), d.float()), torch.from_numpy(out["colbert_fp32"]))
    assert torch.allclose(O.colbert_pairwise_ce_loss_port(q.float(), d.float()), torch.from_numpy(out["pairwise_fp32"]))
    np.savez_compressed(os.path.join(GOLD, "loss_cfg3.npz"), **out)
    print("loss_cfg3 ok")


def loss_smooth():
    """use_smooth_max=True (tau * logsumexp, late_interaction_losses.py:40-44): small cases with all gradients for
    every loss class, and cfg3 (ColbertLoss / ColbertPairwiseCELoss) with fp32 inputs."""
    g = torch.Generator().manual_seed(15)
    rnd = lambda *sh: torch.nn.functional.normalize(torch.randn(*sh, generator=g), dim=-1).bfloat16().float()  # noqa: E731
    q, d, neg = rnd(4, 5, 16), rnd(6, 9, 16), rnd(4, 3, 7, 16)
    q[1, 3:] = 0   # all-zero query rows still add tau * log(N_d) each (but do not count in `lengths`)
    d[0, :2] = 0   # zero document rows take part in the log-sum-exp (exp(0) = 1)
    neg[0, 1, :2] = 0
    out = {"q": q.numpy().copy(), "d": d.numpy().copy(), "neg": neg.numpy().copy()}
    for name, mod, port, kw in (
        ("colbert", ColbertLoss(use_smooth_max=True), O.colbert_loss_port, {}),
        ("colbert_tau05_nonorm", ColbertLoss(use_smooth_max=True, tau=0.05, normalize_scores=False, temperature=0.5),
         O.colbert_loss_port, dict(tau=0.05, normalize_scores=False, temperature=0.5)),
        ("pairwise", ColbertPairwiseCELoss(use_smooth_max=True), O.colbert_pairwise_ce_loss_port, {}),
        ("pairwise_filter", ColbertPairwiseCELoss(use_smooth_max=True, pos_aware_negative_filtering=True),
         O.colbert_pairwise_ce_loss_port, dict(pos_aware_negative_filtering=True)),
    ):
        qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
        loss = mod(qq, dd, offset=1)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_dq"], out[f"{name}_dd"] = loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy()
        assert torch.allclose(port(q, d, offset=1, use_smooth_max=True, **kw), loss.detach(), atol=1e-6), name
    qq, dd = q.clone().requires_grad_(True), d[:4].clone().requires_grad_(True)
    loss = ColbertSigmoidLoss(use_smooth_max=True)(qq, dd)
    loss.backward()
    out["sigmoid_loss"], out["sigmoid_dq"], out["sigmoid_dd"] = loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy()
    for name, mod in (("negce", ColbertNegativeCELoss(use_smooth_max=True)),
                      ("pairneg", ColbertPairwiseNegativeCELoss(use_smooth_max=True, in_batch_term_weight=0.3))):
        qq, dd, nn = q.clone().requires_grad_(True), d.clone().requires_grad_(True), neg.clone().requires_grad_(True)
        loss = mod(qq, dd, nn, offset=1)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_dq"], out[f"{name}_dd"], out[f"{name}_dn"] = (
            loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy(), nn.grad.numpy())
    # cfg3 (BASELINE configs[2]) with the smooth maximum, fp32 inputs = the numerical oracle
    q3, d3, lens = O.cfg3_inputs()
    out["cfg3_q_checksum"], out["cfg3_d_checksum"] = checksum(q3), checksum(d3)
    for name, cls in (("colbert", ColbertLoss), ("pairwise", ColbertPairwiseCELoss)):
        qq, dd = q3.float().requires_grad_(True), d3.float().requires_grad_(True)
        loss = cls(use_smooth_max=True)(qq, dd)
        loss.backward()
        out[f"cfg3_{name}_fp32"] = loss.detach().numpy()
        out[f"cfg3_{name}_dq"] = qq.grad.numpy()
        out[f"cfg3_{name}_dd_first2"] = dd.grad[:2].numpy()
        out[f"cfg3_{name}_dd_rownorm"] = dd.grad.norm(dim=-1).numpy().astype(np.float32)
        print(f"  cfg3 smooth {name}: {float(loss):.5f}")
    np.savez_compressed(os.path.join(GOLD, "loss_smooth.npz"), **out)
    print("loss_smooth ok")


def head_small():
    """Run the reference ColQwen2 (tiny random-init config) and capture the head's input/output."""
    from transformers.models.qwen2_vl import Qwen2VLConfig

    from colpali_engine.models import ColQwen2

    torch.manual_seed(0)
    cfg = Qwen2VLConfig(
        text_config=dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=512, rope_scaling={"type": "mrope", "mrope_section": [8, 12, 12]}),
        vision_config=dict(depth=1, embed_dim=32, hidden_size=256, num_heads=2, patch_size=14),
    )
    out = {}
    for dtype, tag in ((torch.bfloat16, "bf16"), (torch.float32, "fp32")):
        torch.manual_seed(0)
        model = ColQwen2(cfg).to(dtype).eval()
        captured = {}
        model.custom_text_proj.register_forward_hook(lambda m, i, o: captured.__setitem__("h", i[0].detach()))
        ids = torch.randint(0, 500, (3, 24))
        mask = torch.ones(3, 24, dtype=torch.long)
        mask[0, :7] = 0  # left padding (processing_colqwen2.py:43)
        mask[2, :3] = 0
        with torch.no_grad():
            emb = model(input_ids=ids, attention_mask=mask)
        h = captured["h"]
        w, b = model.custom_text_proj.weight.detach(), model.custom_text_proj.bias.detach()
        port = O.head_port(h, w, b, mask)
        assert torch.equal(port, emb), tag
        conv = bits if dtype == torch.bfloat16 else (lambda x: x.numpy())
        out[f"{tag}_h"], out[f"{tag}_w"], out[f"{tag}_b"], out[f"{tag}_out"] = conv(h), conv(w), conv(b), conv(emb)
        out[f"{tag}_mask"] = mask.numpy()
    np.savez_compressed(os.path.join(GOLD, "head_small.npz"), **out)
    print("head_small ok")


def wide_dims():
    """ColQwen3's embedding dim 320 -- scorer (ragged, bf16 and fp32 inputs), ColbertLoss with gradients,
    and the head of a tiny random-init reference ColQwen3 (models/qwen3/colqwen3/modeling_colqwen3.py)."""
    from transformers.models.qwen3_vl import Qwen3VLConfig

    from colpali_engine.models import ColQwen3

    out = {}
    g = torch.Generator().manual_seed(13)
    qlen = [5, 32, 17, 20]
    plen = [1, 16, 255, 256, 257, 600, 1030, 64]
    qs = [torch.nn.functional.normalize(torch.randn(n, 320, generator=g), dim=-1).bfloat16() for n in qlen]
    ps = [torch.nn.functional.normalize(torch.randn(n, 320, generator=g), dim=-1).bfloat16() for n in plen]
    out["s_q"], out["s_qlen"], out["s_p"], out["s_plen"] = bits(torch.cat(qs)), np.array(qlen), bits(torch.cat(ps)), np.array(plen)
    out["s_bf16"] = ref_score(qs, ps).numpy()
    out["s_fp32"] = ref_score([x.float() for x in qs], [x.float() for x in ps]).numpy()
    assert torch.equal(O.score_multi_vector_port(qs, ps), torch.from_numpy(out["s_bf16"]))
    assert torch.equal(O.score_multi_vector_port([x.float() for x in qs], [x.float() for x in ps]), torch.from_numpy(out["s_fp32"]))

    q = torch.nn.functional.normalize(torch.randn(4, 6, 320, generator=g), dim=-1).bfloat16().float()
    d = torch.nn.functional.normalize(torch.randn(6, 40, 320, generator=g), dim=-1).bfloat16().float()
    q[1, 4:] = 0
    d[0, :3] = 0
    out["l_q"], out["l_d"] = q.numpy().copy(), d.numpy().copy()
    for name, mod in (("colbert", ColbertLoss()), ("pairwise", ColbertPairwiseCELoss())):
        qq, dd = q.clone().requires_grad_(True), d.clone().requires_grad_(True)
        loss = mod(qq, dd, offset=1)
        loss.backward()
        out[f"l_{name}_loss"], out[f"l_{name}_dq"], out[f"l_{name}_dd"] = loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy()
    assert torch.allclose(O.colbert_loss_port(q, d, offset=1), torch.from_numpy(out["l_colbert_loss"]))

    torch.manual_seed(0)
    cfg = Qwen3VLConfig(
        text_config=dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, head_dim=64, vocab_size=512,
                         rope_scaling={"rope_type": "default", "mrope_section": [8, 12, 12]}),
        vision_config=dict(depth=1, hidden_size=32, intermediate_size=64, num_heads=2, patch_size=14,
                           out_hidden_size=256, deepstack_visual_indexes=[0]),
    )
    model = ColQwen3(cfg).to(torch.bfloat16).eval()
    captured = {}
    model.custom_text_proj.register_forward_hook(lambda m, i, o: captured.__setitem__("h", i[0].detach()))
    ids = torch.randint(0, 500, (3, 24))
    mask = torch.ones(3, 24, dtype=torch.long)
    mask[0, :7] = 0
    mask[2, :3] = 0
    with torch.no_grad():
        emb = model(input_ids=ids, attention_mask=mask)
    h, w, b = captured["h"], model.custom_text_proj.weight.detach(), model.custom_text_proj.bias.detach()
    assert emb.shape == (3, 24, 320) and torch.equal(O.head_port(h, w, b, mask), emb)
    out["h_h"], out["h_w"], out["h_b"], out["h_out"], out["h_mask"] = bits(h), bits(w), bits(b), bits(emb), mask.numpy()
    np.savez_compressed(os.path.join(GOLD, "wide_dim320.npz"), **out)
    print("wide_dim320 ok")


def bi_small():
    """Row f-4: the six bi-encoder losses (loss + every gradient), score_single_vector and similarity maps, executed by the
    reference on fp32 inputs: B=6, C=12, D=40, n_neg=3; offsets 0 and 6."""
    import importlib.util

    from colpali_engine.loss import bi_encoder_losses as RB

    # the package __init__ imports matplotlib / seaborn (absent here): execute the unmodified module file on its own
    spec = importlib.util.spec_from_file_location(
        "ref_similarity_map_utils", "/root/reference/colpali_engine/interpretability/similarity_map_utils.py")
    smu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(smu)
    get_similarity_maps_from_embeddings = smu.get_similarity_maps_from_embeddings

    g = torch.Generator().manual_seed(11)
    rnd = lambda *sh: torch.nn.functional.normalize(torch.randn(*sh, generator=g), dim=-1)  # noqa: E731
    q, d, neg = rnd(6, 40), rnd(12, 40), rnd(6, 3, 40)
    d[:6] = torch.nn.functional.normalize(q + 0.6 * rnd(6, 40), dim=-1)      # positives at offset 0 ...
    d[6:] = torch.nn.functional.normalize(q + 0.6 * rnd(6, 40), dim=-1)      # ... and at offset 6
    d[3] = torch.nn.functional.normalize(q[1] + 0.05 * rnd(40), dim=-1)      # a negative above 0.95 x the positive (filter)
    out = {"q": q.numpy().copy(), "d": d.numpy().copy(), "neg": neg.numpy().copy()}
    kw_f = dict(pos_aware_negative_filtering=True)
    cases = (
        # name, module, docs, negatives?, offset, port kind, port kwargs
        ("ce", RB.BiEncoderLoss(), d, False, 0, "ce", {}),
        ("ce_off6_filter", RB.BiEncoderLoss(temperature=0.05, **kw_f), d, False, 6, "ce", dict(temperature=0.05, **kw_f)),
        ("paired", RB.BiPairedEncoderLoss(), d[:6], False, 0, "paired", {}),
        ("paired_filter", RB.BiPairedEncoderLoss(temperature=0.1, **kw_f), d[:6], False, 0, "paired", dict(temperature=0.1, **kw_f)),
        ("pairwise", RB.BiPairwiseCELoss(), d, False, 0, "pairwise", {}),
        ("pairwise_filter_t1", RB.BiPairwiseCELoss(temperature=1.0, **kw_f), d, False, 0, "pairwise", dict(temperature=1.0, **kw_f)),
        ("sigmoid", RB.BiSigmoidLoss(), d[:6], False, 0, "sigmoid", {}),
        ("sigmoid_off6_filter", RB.BiSigmoidLoss(temperature=0.5, **kw_f), d, False, 6, "sigmoid", dict(temperature=0.5, **kw_f)),
        ("negce", RB.BiNegativeCELoss(), d, True, 6, "negce", {}),
        ("negce_w0", RB.BiNegativeCELoss(in_batch_term_weight=0.0), d, True, 0, "negce", dict(in_batch_term_weight=0.0)),
        ("negce_filter", RB.BiNegativeCELoss(temperature=0.1, in_batch_term_weight=0.3, **kw_f), d, True, 0, "negce",
         dict(temperature=0.1, in_batch_term_weight=0.3, **kw_f)),
        ("pairneg", RB.BiPairwiseNegativeCELoss(), d, True, 0, "pairneg", {}),
        ("pairneg_off6", RB.BiPairwiseNegativeCELoss(temperature=0.5, in_batch_term_weight=0.7), d, True, 6, "pairneg",
         dict(temperature=0.5, in_batch_term_weight=0.7)),
    )
    for name, mod, docs, with_neg, off, kind, kw in cases:
        qq, dd = q.clone().requires_grad_(True), docs.clone().requires_grad_(True)
        nn = neg.clone().requires_grad_(True) if with_neg else None
        loss = mod(qq, dd, nn, offset=off) if with_neg else mod(qq, dd, offset=off)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_dq"], out[f"{name}_dd"] = loss.detach().numpy(), qq.grad.numpy(), dd.grad.numpy()
        if with_neg:
            out[f"{name}_dn"] = nn.grad.numpy()
        port = O.bi_loss_port(kind, q, docs, neg if with_neg else None, offset=off, **kw)
        assert torch.allclose(port, loss.detach(), atol=1e-6, rtol=1e-6), (name, float(port), float(loss))
    # score_single_vector (processing_utils.py:103-130), fp32 and bf16 operands
    out["single_f32"] = RefProc.score_single_vector(q, d, device="cpu").numpy()
    out["single_bf16"] = RefProc.score_single_vector(list(q.bfloat16()), list(d.bfloat16()), device="cpu").numpy()
    assert torch.equal(O.score_single_vector_port(q, d), torch.from_numpy(out["single_f32"]))
    # similarity maps (similarity_map_utils.py:9-56): 2 images, 3 x 4 and 2 x 5 patch grids inside 14 / 12 tokens
    img, qe = rnd(2, 14, 16), rnd(2, 5, 16)
    mask = torch.zeros(2, 14, dtype=torch.bool)
    mask[0, 1:13] = True
    mask[1, [0, 2, 3, 4, 6, 7, 9, 10, 11, 13]] = True
    n_patches = [(3, 4), (2, 5)]
    maps = get_similarity_maps_from_embeddings(img, qe, n_patches, mask)
    port = O.similarity_maps_port(img, qe, n_patches, mask)
    out["map_img"], out["map_q"], out["map_mask"] = img.numpy().copy(), qe.numpy().copy(), mask.numpy().copy()
    for k, m in enumerate(maps):
        assert m.shape == (5, *n_patches[k]) and torch.allclose(port[k], m, atol=1e-6)
        out[f"map_{k}"] = m.numpy()
    np.savez_compressed(os.path.join(GOLD, "bi_small.npz"), **out)
    print("bi_small ok")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = set(sys.argv[1:]) or {"scorer_small", "cfg1", "cfg2", "loss_small", "loss_neg_small", "loss_cfg3", "loss_smooth",
                                  "head_small", "wide", "bi"}
    if "scorer_small" in which:
        scorer_small()
    if "cfg1" in which:
        scorer_cfg("cfg1", *O.cfg1_inputs())
    if "cfg2" in which:
        scorer_cfg("cfg2", *O.cfg2_inputs())
    if "loss_small" in which:
        loss_small()
    if "loss_neg_small" in which:
        loss_neg_small()
    if "loss_cfg3" in which:
        loss_cfg3()
    if "loss_smooth" in which:
        loss_smooth()
    if "head_small" in which:
        head_small()
    if "wide" in which:
        wide_dims()
    if "bi" in which:
        bi_small()
