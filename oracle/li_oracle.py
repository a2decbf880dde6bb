"""CPU oracle for the late-interaction hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker or the timed CPU arm.  Nothing under
``colpali_b200/`` imports it; the product path has no CPU fallback.

What it restates (reference = illuin-tech/colpali @ 9be8f19, paths relative to /root/reference):

* ``score_multi_vector_port``  <- colpali_engine/utils/processing_utils.py:132-187
  (batching :170/:175, zero padding :172/:176-178, einsum/max/sum :179, fp32 cast :186).
* ``colbert_scores_port`` / ``colbert_loss_port`` / ``colbert_pairwise_ce_loss_port``
  <- colpali_engine/loss/late_interaction_losses.py:73-107 (aggregate, filter), :140-164, :284-313;
  ``colbert_negative_ce_loss_port`` :215-252, ``colbert_pairwise_negative_ce_loss_port`` :361-398,
  ``colbert_sigmoid_loss_port`` :431-465.
* ``bi_loss_port`` <- colpali_engine/loss/bi_encoder_losses.py:64-418 (the six bi-encoder losses, one closed form),
  ``similarity_maps_port`` <- colpali_engine/interpretability/similarity_map_utils.py:9-56,
  ``score_single_vector_port`` <- colpali_engine/utils/processing_utils.py:103-130.
* ``head_port`` <- colpali_engine/models/qwen2/colqwen2/modeling_colqwen2.py:65-74
  (and modernvbert's clamp variant, models/modernvbert/colvbert/modeling_colmodernvbert.py:59).
* ``maxsim_f64`` -- an independent numpy float64 evaluation of sum_n max_s <q_n, d_s> used to
  cross-check the port itself.

The arithmetic of the reference lives in PyTorch ATen (torch.einsum -> bmm, max, sum, ...; torch is a
third-party dependency pinned ``torch>=2.2.0,<2.11.0`` in pyproject.toml:41, 2.11.0+cu128 here), so the
port calls the same ATen CPU kernels in the same order and dtype; it is pinned against outputs of the
reference itself run in the build container (``oracle/make_golden.py`` -> ``tests/golden/``).
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F  # noqa: N812

TensorOrList = Union[torch.Tensor, List[torch.Tensor]]


# ---------------------------------------------------------------------------------------------
# scorer
# ---------------------------------------------------------------------------------------------
def score_multi_vector_port(qs: TensorOrList, ps: TensorOrList, batch_size: int = 128,
                            device: Union[str, torch.device] = "cpu") -> torch.Tensor:
    """processing_utils.py:132-187 restated; same padding groups, same dtype path, CPU fp32 out."""
    if len(qs) == 0:
        raise ValueError("No queries provided")
    if len(ps) == 0:
        raise ValueError("No passages provided")
    rows = []
    for i in range(0, len(qs), batch_size):
        qb = torch.nn.utils.rnn.pad_sequence(list(qs[i : i + batch_size]), batch_first=True, padding_value=0).to(device)
        cols = []
        for j in range(0, len(ps), batch_size):
            pb = torch.nn.utils.rnn.pad_sequence(list(ps[j : j + batch_size]), batch_first=True, padding_value=0).to(device)
            sim = torch.einsum("bnd,csd->bcns", qb, pb)  # [b, c, n, s] in the input dtype
            cols.append(sim.max(dim=3)[0].sum(dim=2))
        rows.append(torch.cat(cols, dim=1).cpu())
    scores = torch.cat(rows, dim=0)
    assert scores.shape[0] == len(qs)
    return scores.to(torch.float32)


def maxsim_f64(qs: Sequence, ps: Sequence, floors: Optional[Sequence[float]] = None) -> np.ndarray:
    """Independent float64 numpy evaluation (no batching, no padding unless ``floors`` says so)."""
    out = np.zeros((len(qs), len(ps)), dtype=np.float64)
    for j, p in enumerate(ps):
        pj = np.asarray(p.detach().float().cpu().numpy() if isinstance(p, torch.Tensor) else p, dtype=np.float64)
        for i, q in enumerate(qs):
            qi = np.asarray(q.detach().float().cpu().numpy() if isinstance(q, torch.Tensor) else q, dtype=np.float64)
            if pj.shape[0] == 0:
                mx = np.full(qi.shape[0], -np.inf)
            else:
                mx = (qi @ pj.T).max(axis=1)
            if floors is not None:
                mx = np.maximum(mx, floors[j])
            out[i, j] = mx.sum()
    return out


def reference_floors(lens: Sequence[int], batch_size: int = 128) -> List[float]:
    """0.0 where processing_utils.py:176-178 would zero-pad the document inside its batch, else -inf."""
    fl = []
    for j in range(0, len(lens), batch_size):
        chunk = list(lens[j : j + batch_size])
        mx = max(chunk)
        fl += [0.0 if n < mx else -math.inf for n in chunk]
    return fl


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------
def colbert_scores_port(q: torch.Tensor, d: torch.Tensor, normalize_scores: bool = True,
                        use_smooth_max: bool = False, tau: float = 0.1) -> torch.Tensor:
    """late_interaction_losses.py:152-156 (+ :40-44, :73-91): [B, C] (optionally length-normalised) scores."""
    lengths = (q[:, :, 0] != 0).sum(dim=1)
    raw = torch.einsum("bnd,csd->bcns", q, d)
    if use_smooth_max:
        scores = (tau * torch.logsumexp(raw / tau, dim=3)).sum(dim=2)
    else:
        scores = raw.amax(dim=3).sum(dim=2)
    if normalize_scores:
        scores = scores / lengths.unsqueeze(1)
    return scores


def _filter_high_negatives_port(scores: torch.Tensor, pos_idx: torch.Tensor, thr: float, factor: float) -> torch.Tensor:
    """late_interaction_losses.py:93-107 (out of place)."""
    b = scores.size(0)
    idx = torch.arange(b, device=scores.device)
    pos = scores[idx, pos_idx]
    mask = scores > thr * pos.unsqueeze(1)
    mask[idx, pos_idx] = False
    return torch.where(mask, scores * factor, scores)


def colbert_loss_port(q: torch.Tensor, d: torch.Tensor, offset: int = 0, temperature: float = 0.02,
                      normalize_scores: bool = True, use_smooth_max: bool = False,
                      pos_aware_negative_filtering: bool = False, tau: float = 0.1,
                      filter_threshold: float = 0.95, filter_factor: float = 0.5) -> torch.Tensor:
    """ColbertLoss.forward, late_interaction_losses.py:140-164."""
    scores = colbert_scores_port(q, d, normalize_scores, use_smooth_max, tau)
    b = scores.size(0)
    pos_idx = torch.arange(b, device=scores.device) + offset
    if pos_aware_negative_filtering:
        scores = _filter_high_negatives_port(scores, pos_idx, filter_threshold, filter_factor)
    return F.cross_entropy(scores / temperature, pos_idx)


def colbert_pairwise_ce_loss_port(q: torch.Tensor, d: torch.Tensor, offset: int = 0, temperature: float = 1.0,
                                  normalize_scores: bool = True, use_smooth_max: bool = False,
                                  pos_aware_negative_filtering: bool = False, tau: float = 0.1,
                                  filter_threshold: float = 0.95, filter_factor: float = 0.5) -> torch.Tensor:
    """ColbertPairwiseCELoss.forward, late_interaction_losses.py:284-313."""
    scores = colbert_scores_port(q, d, normalize_scores, use_smooth_max, tau)
    b = scores.size(0)
    pos_idx = torch.arange(b, device=scores.device) + offset
    if pos_aware_negative_filtering:
        scores = _filter_high_negatives_port(scores, pos_idx, filter_threshold, filter_factor)
    pos = scores.diagonal(offset=offset)
    top2 = scores.topk(2, dim=1).values
    neg = torch.where(top2[:, 0] == pos, top2[:, 1], top2[:, 0])
    return F.softplus((neg - pos) / temperature).mean()


def _paired_neg_term(q, d, neg, offset, temperature, normalize_scores):
    """softplus((neg - pos) / T).mean() over every query's own negatives, late_interaction_losses.py:234-246 / :380-392."""
    lengths = (q[:, :, 0] != 0).sum(dim=1)
    pos = torch.einsum("bnd,bsd->bns", q, d[offset : offset + neg.size(0)]).amax(dim=2).sum(dim=1)
    negs = torch.einsum("bnd,blsd->blns", q, neg).amax(dim=3).sum(dim=2)
    if normalize_scores:
        pos = pos / lengths
        negs = negs / lengths.unsqueeze(1)
    return F.softplus((negs - pos.unsqueeze(1)) / temperature).mean()


def colbert_negative_ce_loss_port(q, d, neg, offset: int = 0, temperature: float = 0.02, normalize_scores: bool = True,
                                  pos_aware_negative_filtering: bool = False, in_batch_term_weight: float = 0.5,
                                  filter_threshold: float = 0.95, filter_factor: float = 0.5) -> torch.Tensor:
    """ColbertNegativeCELoss.forward, late_interaction_losses.py:215-252."""
    loss = _paired_neg_term(q, d, neg, offset, temperature, normalize_scores)
    if in_batch_term_weight > 0:
        ib = colbert_loss_port(q, d, offset, temperature, normalize_scores, False, pos_aware_negative_filtering,
                               0.1, filter_threshold, filter_factor)
        loss = loss * (1 - in_batch_term_weight) + ib * in_batch_term_weight
    return loss


def colbert_pairwise_negative_ce_loss_port(q, d, neg, offset: int = 0, temperature: float = 0.02,
                                           normalize_scores: bool = True, pos_aware_negative_filtering: bool = False,
                                           in_batch_term_weight: float = 0.5, filter_threshold: float = 0.95,
                                           filter_factor: float = 0.5) -> torch.Tensor:
    """ColbertPairwiseNegativeCELoss.forward, late_interaction_losses.py:361-398."""
    loss = _paired_neg_term(q, d, neg, offset, temperature, normalize_scores)
    if in_batch_term_weight > 0:
        ib = colbert_pairwise_ce_loss_port(q, d, offset, temperature, normalize_scores, False,
                                           pos_aware_negative_filtering, 0.1, filter_threshold, filter_factor)
        loss = loss * (1 - in_batch_term_weight) + ib * in_batch_term_weight
    return loss


def colbert_sigmoid_loss_port(q, d, offset: int = 0, temperature: float = 0.02, normalize_scores: bool = True,
                              pos_aware_negative_filtering: bool = False, filter_threshold: float = 0.95,
                              filter_factor: float = 0.5) -> torch.Tensor:
    """ColbertSigmoidLoss.forward, late_interaction_losses.py:431-465."""
    scores = colbert_scores_port(q, d, normalize_scores)
    b = scores.size(0)
    pos_idx = torch.arange(b, device=scores.device) + offset
    if pos_aware_negative_filtering:
        scores = _filter_high_negatives_port(scores, pos_idx, filter_threshold, filter_factor)
    pos_mask = -torch.ones(b * b, device=scores.device)
    pos_mask[pos_idx * (b + 1)] = 1.0
    return F.softplus(-(scores.reshape(-1) / temperature) * pos_mask).mean()


# ---------------------------------------------------------------------------------------------
# single-vector (bi-encoder) scorer, losses, similarity maps                      
# ---------------------------------------------------------------------------------------------
def score_single_vector_port(qs: torch.Tensor, ps: torch.Tensor) -> torch.Tensor:
    """processing_utils.py:126-129."""
    return torch.einsum("bd,cd->bc", qs, ps).to(torch.float32)


def bi_loss_port(kind: str, q: torch.Tensor, d: torch.Tensor, neg: Optional[torch.Tensor] = None, offset: int = 0,
                 temperature: float = 0.02, pos_aware_negative_filtering: bool = False, in_batch_term_weight: float = 0.5,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5) -> torch.Tensor:
    """The bi-encoder losses of bi_encoder_losses.py written as ONE closed form over the [B, C] score matrix (what the loss
    kernel evaluates) instead of the reference's module-by-module code:
      kind "ce" BiEncoderLoss :103-113, "paired" BiPairedEncoderLoss :157-168, "pairwise" BiPairwiseCELoss :290-302
      (positives = diagonal, offset ignored), "sigmoid" BiSigmoidLoss :396-418 (all B x C pairs, +1 at b + offset),
      "negce" BiNegativeCELoss :231-248, "pairneg" BiPairwiseNegativeCELoss :350-358."""
    t = temperature
    scores = torch.einsum("bd,cd->bc", q, d)
    b, c = scores.shape
    idx = torch.arange(b)
    ib_offset = 0 if kind in ("pairwise", "pairneg") else offset
    pos_idx = idx + ib_offset
    filt = pos_aware_negative_filtering and kind != "pairneg"
    if filt:
        scores = _filter_high_negatives_port(scores, pos_idx, filter_threshold, filter_factor)

    def pairwise(sc):
        pos = sc.diagonal()
        top2 = sc.topk(2, dim=1).values
        hard = torch.where(top2[:, 0] == pos, top2[:, 1], top2[:, 0])
        return F.softplus((hard - pos) / t).mean()

    if kind == "ce":
        return F.cross_entropy(scores / t, pos_idx)
    if kind == "paired":
        return (F.cross_entropy(scores / t, pos_idx) + F.cross_entropy(scores.T / t, idx)) / 2.0
    if kind == "pairwise":
        return pairwise(scores)
    if kind == "sigmoid":
        labels = -torch.ones(b, c)
        labels[idx, pos_idx] = 1.0
        return F.softplus(-(scores / t) * labels).mean()
    pos = (q * d[offset : offset + b]).sum(dim=1)
    negs = torch.einsum("bd,bnd->bn", q, neg)
    loss = F.softplus((negs - pos.unsqueeze(1)) / t).mean()
    if in_batch_term_weight > 0:
        ib = F.cross_entropy(scores / t, pos_idx) if kind == "negce" else pairwise(scores)
        loss = loss * (1 - in_batch_term_weight) + ib * in_batch_term_weight
    return loss


def similarity_maps_port(image_embeddings: torch.Tensor, query_embeddings: torch.Tensor, n_patches, image_mask: torch.Tensor):
    """similarity_map_utils.py:29-56 with plain indexing instead of einops: map[n, i, j] = <q_n, patch (row j, column i)>."""
    maps = []
    for k in range(image_embeddings.size(0)):
        w, h = n_patches[k] if isinstance(n_patches, list) else n_patches
        patches = image_embeddings[k][image_mask[k]]           # (h * w, dim), row-major over (h, w)
        grid = patches.view(h, w, -1).permute(1, 0, 2)         # (w, h, dim)
        maps.append(torch.einsum("nk,ijk->nij", query_embeddings[k], grid))
    return maps


# ---------------------------------------------------------------------------------------------
# projection head
# ---------------------------------------------------------------------------------------------
def head_port(h: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], attention_mask: torch.Tensor,
              image_mask: Optional[torch.Tensor] = None, clamp_norm: bool = False) -> torch.Tensor:
    """custom_text_proj -> / L2 norm -> * attention_mask [-> * image_mask], modeling_colqwen2.py:65-74."""
    proj = F.linear(h, weight, bias)
    norm = proj.norm(dim=-1, keepdim=True)
    if clamp_norm:  # modeling_colmodernvbert.py:59
        norm = norm.clamp_min(1e-12)
    proj = proj / norm
    proj = proj * attention_mask.unsqueeze(-1)
    if image_mask is not None:
        proj = proj * image_mask.unsqueeze(-1)
    return proj


# ---------------------------------------------------------------------------------------------
# deterministic inputs shared by golden generation, tests and bench
# ---------------------------------------------------------------------------------------------
def unit_rows(shape, seed: int, dtype=torch.bfloat16) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return F.normalize(torch.randn(*shape, generator=g), dim=-1).to(dtype)


def cfg1_inputs():
    return unit_rows((4, 32, 128), 0), unit_rows((16, 256, 128), 1)


def cfg2_inputs(n_docs: int = 1000):
    return unit_rows((32, 32, 128), 0), unit_rows((n_docs, 1030, 128), 1)


def cfg3_inputs(batch: int = 64, n_q: int = 32, max_len: int = 1030, min_len: int = 768, dim: int = 128):
    """In-batch-negatives batch: ragged docs LEFT-padded with zero rows (contrastive_trainer.py:145-147),
    positives planted so the loss is non-trivial.  Returns (q [B,Nq,D], d [B,L,D], lens [B]) in bf16."""
    g = torch.Generator().manual_seed(1)
    lens = torch.randint(min_len, max_len + 1, (batch,), generator=g)
    q = unit_rows((batch, n_q, dim), 0, torch.float32)
    d = unit_rows((batch, max_len, dim), 2, torch.float32)
    noise = torch.randn(batch, n_q, dim, generator=torch.Generator().manual_seed(3))  # |noise| ~ sqrt(dim)
    d[:, -n_q:] = F.normalize(q + 0.5 * noise, dim=-1)
    for j in range(batch):
        d[j, : max_len - int(lens[j])] = 0
    return q.bfloat16(), d.bfloat16(), lens
