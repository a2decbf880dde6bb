"""ColBERT in-batch-negative losses on B200 behind the reference's loss-module API.

Mirrors ``colpali_engine/loss/late_interaction_losses.py`` (reference @ 9be8f19):

* ``ColbertModule``           :6-107   (hyper-parameters only; the reference's eager helper methods ``_aggregate`` /
                                        ``_smooth_max`` / ``_filter_high_negatives`` live inside the kernels here)
* ``ColbertLoss``             :110-164 (InfoNCE over in-batch documents)
* ``ColbertPairwiseCELoss``   :255-313 (softplus(hardest in-batch negative - positive))
* ``ColbertNegativeCELoss``   :167-252, ``ColbertPairwiseNegativeCELoss`` :316-398 (explicit negatives)
* ``ColbertSigmoidLoss``      :401-465

Same constructor keyword arguments, same ``forward(query_embeddings, doc_embeddings, offset=0)``.  The
``einsum("bnd,csd->bcns")`` / ``amax`` (or ``tau * logsumexp`` with ``use_smooth_max=True``, :40-44) / ``sum`` of
:153-154 and everything after it run as

1. ONE launch of the fused sm_100a MaxSim kernel (csrc/maxsim_sm100.cu; csrc/maxsim_kpipe_sm100.cu for embedding dims
   above 128) whose last CTA turns the ``[B, C]`` sums into the scalar loss *and* its gradient with respect to the
   sums (csrc/loss_body.cuh).  For the backward it records, per (document, query token), either the index of the
   winning document token (int32) or the smooth maximum (fp32) -- ``[C, B*N_q]`` instead of the reference's saved
   ``[B, C, N_q, N_d]`` similarity tensor.  (Queries longer than 32 tokens and the explicit-negative losses, which need
   two score matrices, run the loss as its own small kernel, csrc/loss_sm100.cu.)
2. in backward, hard max: a gather kernel for ``dQ`` and a per-document counting-sort kernel that writes every row of
   ``dD`` once (csrc/loss_sm100.cu); smooth max: two kernels that recompute the similarity tiles on the tensor cores
   and apply the softmax weights (csrc/smooth_bwd_sm100.cu, embedding dim 128).

Differences from the reference that a caller can observe, all documented in DESIGN.md:
the loss is returned in fp32 whatever the embedding dtype (the reference returns the embedding dtype);
embeddings are contracted in bf16 with fp32 accumulation; where several document tokens tie for the maximum the
gradient goes to the first one (``amax`` splits it evenly) -- this only differs on all-zero (padding) rows, whose
gradients the model masks anyway.
"""

from __future__ import annotations

import ctypes

import torch

from . import _lib
from .scoring import DocBank, QueryBlock, _on_device, launch_maxsim


_DONE_COUNTERS: dict = {}
_DONE_WORDS = 4096


def _done_counter(dev: torch.device) -> torch.Tensor:
    """Counter workspace of the fused loss (``cpb_maxsim_args.d_done_counter``: completion counters per query-tile group
    and their partial sums, ``CPB_LOSS_WORKSPACE_WORDS`` = 4096 words: up to 1023 groups = 8184 queries), one per
    (device, stream); zero once, the kernel leaves the counters zero."""
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    t = _DONE_COUNTERS.get(key)
    if t is None:
        t = torch.zeros(_DONE_WORDS, dtype=torch.int32, device=dev)
        _DONE_COUNTERS[key] = t
    return t


def _loss_desc(mode, temperature, normalize, filt, thr, factor, offset, loss, grad, bounds=None, neg_scores=None,
               n_neg=0, weight=1.0, grad_neg=None) -> _lib.LossDesc:
    d = _lib.LossDesc()
    d.mode, d.normalize_scores, d.pos_aware_negative_filtering, d.offset = int(mode), int(normalize), int(filt), int(offset)
    d.temperature, d.filter_threshold, d.filter_factor = float(temperature), float(thr), float(factor)
    d.d_neg_scores = neg_scores.data_ptr() if neg_scores is not None else None
    d.n_neg, d.in_batch_term_weight = int(n_neg), float(weight)
    d.d_loss = loss.data_ptr()
    d.d_grad_scores = grad.data_ptr() if grad is not None else None
    d.d_grad_neg_scores = grad_neg.data_ptr() if grad_neg is not None else None
    d.d_bounds = bounds.data_ptr() if bounds is not None else None
    return d


def _maxsim_for_loss(qb: QueryBlock, bank: DocBank, need_grad: bool, smooth_tau: float, nq_real: int,
                     loss_desc=None, wait=None):
    """One fused MaxSim launch; returns (scores, aux) with aux = int32 argmax (hard max) or fp32 lse (smooth max), the
    only per-(document, query row) state the backward needs.  With ``loss_desc`` the last CTA also emits the loss."""
    dev = bank.device
    scores = torch.empty(qb.n, bank.n_docs, dtype=torch.float32, device=dev)
    aux = None
    if need_grad:
        aux = torch.empty(bank.n_docs, qb.n * qb.nq_pad, dtype=torch.float32 if smooth_tau > 0 else torch.int32, device=dev)
    launch_maxsim(qb, bank, scores=scores, argmax=aux if smooth_tau == 0 else None, lse=aux if smooth_tau > 0 else None,
                  smooth_tau=smooth_tau, nq_real=nq_real, loss=loss_desc,
                  done_counter=_done_counter(dev) if loss_desc is not None else None, wait=wait)
    return scores, aux


def _maxsim_backward(g, go, aux, smooth_tau, nq_real, q_flat, b, nq_pad, bank_flat, bank_start, bank_len, n_docs,
                     max_len, need_dq, need_dd, dd_doc_base=None, bf16_out=False):
    """dq [b * nq_pad, dim], dd [rows, dim] (fp32, fully written by the kernels) for one score matrix.  ``bf16_out``
    (hard max only): the kernels round the rows to bf16 themselves, saving the two cast kernels of a bf16 model.  With
    ``dd_doc_base`` (int64 device tensor of per-document peer addresses) the document gradients are added straight into
    their owner ranks' accumulators instead (multi-GPU exchange) and ``dd`` is None."""
    dev = q_flat.device
    dim = q_flat.shape[1]
    bf16_out = bool(bf16_out) and not smooth_tau > 0 and dd_doc_base is None
    gdt = torch.bfloat16 if bf16_out else torch.float32
    dq = torch.empty(b * nq_pad, dim, dtype=gdt, device=dev) if need_dq else None
    dd = torch.empty(bank_flat.shape[0], dim, dtype=gdt, device=dev) if (need_dd and dd_doc_base is None) else None
    if not (need_dq or need_dd):
        return None, None
    a = _lib.MaxSimBwdArgs()
    a.flags = _lib.CPB_FLAG_CONTIGUOUS  # the banks of the losses are dense [C, L, D] tensors
    if bf16_out:
        a.flags |= _lib.CPB_FLAG_GRAD_BF16
    a.d_grad_scores, a.d_grad_out = g.data_ptr(), go.data_ptr()
    if smooth_tau > 0:
        a.d_lse, a.smooth_tau = aux.data_ptr(), float(smooth_tau)
    else:
        a.d_argmax = aux.data_ptr()
    a.d_q, a.n_queries, a.nq_pad, a.nq_real, a.dim = q_flat.data_ptr(), b, nq_pad, int(nq_real), dim
    a.d_docs, a.doc_rows = bank_flat.data_ptr(), bank_flat.shape[0]
    a.d_doc_start, a.d_doc_len, a.n_docs, a.max_doc_len = bank_start.data_ptr(), bank_len.data_ptr(), n_docs, int(max_len)
    a.d_dq = dq.data_ptr() if need_dq else None
    a.d_dd = dd.data_ptr() if dd is not None else None
    a.d_dd_doc_base = dd_doc_base.data_ptr() if (need_dd and dd_doc_base is not None) else None
    with _on_device(dev):
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.load().cpb_maxsim_bwd_launch(ctypes.byref(a))
    _lib.check(rc, "cpb_maxsim_bwd_launch")
    # hard max: dQ and dD blocks share one launch; smooth max: one recompute kernel each
    _lib.count_launches(int(need_dq) + int(need_dd) if smooth_tau > 0 else 1)
    return dq, dd


def _check_inputs(q, d, neg=None):
    if q.dim() != 3 or d.dim() != 3 or (neg is not None and neg.dim() != 4):
        raise ValueError("expected [B, N_q, D], [C, N_d, D] (and [B, n_neg, N_neg, D]) embeddings, got "
                         f"{tuple(q.shape)} / {tuple(d.shape)}" + (f" / {tuple(neg.shape)}" if neg is not None else ""))
    dev = q.device
    if dev.type != "cuda" or d.device != dev or (neg is not None and neg.device != dev):
        raise _lib.ColpaliB200Error("colpali_b200 losses need all embeddings on the same CUDA device")
    return dev


class _InBatchLossFn(torch.autograd.Function):
    """ColbertLoss / ColbertPairwiseCELoss / ColbertSigmoidLoss: ONE kernel launch forward -- the fused MaxSim kernel
    whose last CTA turns the [B, C] score matrix into the loss and d loss / d scores (queries of up to 32 tokens; longer
    queries add the segment-sum and loss kernels)."""

    @staticmethod
    def forward(ctx, q, d, offset, mode, temperature, normalize, filt, thr, factor, bounds_out, smooth_tau):
        dev = _check_inputs(q, d)
        qb = QueryBlock(q.detach(), dev)
        bank = DocBank.from_passages(d.detach(), dev)  # dense [C, L, D]: zero (padding) rows are ordinary tokens,
        # they score exactly 0 and take part in the max / log-sum-exp, as in the reference
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        b, c = qb.n, bank.n_docs
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g = torch.empty(b, c, dtype=torch.float32, device=dev) if need_grad else None
        nq_real = int(q.shape[1])
        if qb.nq_pad == 32:
            desc = _loss_desc(mode, temperature, normalize, filt, thr, factor, offset, loss, g, bounds_out)
            _, aux = _maxsim_for_loss(qb, bank, need_grad, smooth_tau, nq_real, loss_desc=desc)
        else:
            scores, aux = _maxsim_for_loss(qb, bank, need_grad, smooth_tau, nq_real)
            desc = _loss_desc(mode, temperature, normalize, filt, thr, factor, offset, loss, g, bounds_out)
            with _on_device(dev):
                rc = _lib.load().cpb_colbert_loss_launch(ctypes.byref(desc), scores.data_ptr(), qb.flat.data_ptr(), b,
                                                         qb.nq_pad, c, qb.flat.shape[1],
                                                         torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "cpb_colbert_loss_launch")
            _lib.count_launches(1)
        if need_grad:
            ctx.save_for_backward(qb.flat, bank.flat, bank.start, bank.length, aux, g)
            ctx.meta = (b, qb.nq_pad, nq_real, c, bank.max_len, float(smooth_tau), tuple(q.shape), tuple(d.shape),
                        q.dtype, d.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        q_flat, d_flat, d_start, d_len, aux, g = ctx.saved_tensors
        b, nq_pad, nq_real, c, max_len, smooth_tau, q_shape, d_shape, q_dtype, d_dtype = ctx.meta
        want_q, want_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dim = q_flat.shape[1]  # padded embedding dim (128, or 192 / 256 / 320 for wide models)
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        both_bf16 = (q_dtype == torch.bfloat16 or not want_q) and (d_dtype == torch.bfloat16 or not want_d)
        dq, dd = _maxsim_backward(g, go, aux, smooth_tau, nq_real, q_flat, b, nq_pad, d_flat, d_start, d_len, c, max_len,
                                  want_q, want_d, bf16_out=both_bf16)
        grad_q = grad_d = None
        if want_q:
            grad_q = dq.view(b, nq_pad, dim)[:, : q_shape[1], : q_shape[2]].to(q_dtype)
        if want_d:
            grad_d = dd.view(d_shape[0], d_shape[1], dim)[:, :, : d_shape[2]].to(d_dtype)
        return (grad_q, grad_d) + (None,) * 9


class _NegLossFn(torch.autograd.Function):
    """Explicit-negative losses: one dense MaxSim launch against the (gathered) positives, one against the flattened
    negatives of all queries (only the block diagonal is used), one loss kernel for both score matrices."""

    @staticmethod
    def forward(ctx, q, d, neg, offset, inner_mode, temperature, normalize, filt, thr, factor, weight, smooth_tau):
        dev = _check_inputs(q, d, neg)
        if neg.shape[0] != q.shape[0]:
            raise ValueError(f"{neg.shape[0]} negative groups for {q.shape[0]} queries")
        qb = QueryBlock(q.detach(), dev)
        bank = DocBank.from_passages(d.detach(), dev)
        b, n_neg = neg.shape[0], neg.shape[1]
        nbank = DocBank.from_passages(neg.detach().reshape(b * n_neg, neg.shape[2], neg.shape[3]), dev)
        need_grad = any(ctx.needs_input_grad[:3])
        nq_real = int(q.shape[1])
        s_pos, aux_pos = _maxsim_for_loss(qb, bank, need_grad, smooth_tau, nq_real)
        s_neg, aux_neg = _maxsim_for_loss(qb, nbank, need_grad, smooth_tau, nq_real)
        c = bank.n_docs
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g_pos = torch.empty(b, c, dtype=torch.float32, device=dev) if need_grad else None
        g_neg = torch.empty(b, b * n_neg, dtype=torch.float32, device=dev) if need_grad else None
        desc = _loss_desc(inner_mode, temperature, normalize, filt, thr, factor, offset, loss, g_pos, None, s_neg, n_neg,
                          weight, g_neg)
        with _on_device(dev):
            rc = _lib.load().cpb_colbert_loss_launch(ctypes.byref(desc), s_pos.data_ptr(), qb.flat.data_ptr(), b, qb.nq_pad,
                                                     c, qb.flat.shape[1], torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "cpb_colbert_loss_launch")
        _lib.count_launches(1)
        if need_grad:
            ctx.save_for_backward(qb.flat, bank.flat, bank.start, bank.length, aux_pos, g_pos, nbank.flat, nbank.start,
                                  nbank.length, aux_neg, g_neg)
            ctx.meta = (b, qb.nq_pad, nq_real, c, b * n_neg, bank.max_len, nbank.max_len, float(smooth_tau), tuple(q.shape),
                        tuple(d.shape), tuple(neg.shape), q.dtype, d.dtype, neg.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        (q_flat, d_flat, d_start, d_len, aux_pos, g_pos, n_flat, n_start, n_len, aux_neg, g_neg) = ctx.saved_tensors
        (b, nq_pad, nq_real, c, cn, max_len, nmax_len, smooth_tau, q_shape, d_shape, n_shape, q_dtype, d_dtype,
         n_dtype) = ctx.meta
        want_q, want_d, want_n = ctx.needs_input_grad[:3]
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        dim = q_flat.shape[1]
        dq1, dd = _maxsim_backward(g_pos, go, aux_pos, smooth_tau, nq_real, q_flat, b, nq_pad, d_flat, d_start, d_len, c,
                                   max_len, want_q, want_d)
        dq2, dn = _maxsim_backward(g_neg, go, aux_neg, smooth_tau, nq_real, q_flat, b, nq_pad, n_flat, n_start, n_len, cn,
                                   nmax_len, want_q, want_n)
        grad_q = grad_d = grad_n = None
        if want_q:
            grad_q = (dq1 + dq2).view(b, nq_pad, dim)[:, : q_shape[1], : q_shape[2]].to(q_dtype)
        if want_d:
            grad_d = dd.view(d_shape[0], d_shape[1], dim)[:, :, : d_shape[2]].to(d_dtype)
        if want_n:
            grad_n = dn.view(n_shape[0], n_shape[1], n_shape[2], dim)[..., : n_shape[3]].to(n_dtype)
        return (grad_q, grad_d, grad_n) + (None,) * 9


class ColbertModule(torch.nn.Module):
    """late_interaction_losses.py:6-31 -- the hyper-parameters shared by the ColBERT losses and the fused dispatch."""

    def __init__(self, max_batch_size: int = 1024, tau: float = 0.1, norm_tol: float = 1e-3,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__()
        self.register_buffer("idx_buffer", torch.arange(max_batch_size), persistent=False)
        self.tau = tau
        self.norm_tol = norm_tol
        self.filter_threshold = filter_threshold
        self.filter_factor = filter_factor
        # set True to reproduce the reference's "Scores out of bounds after normalization" print (:64-70);
        # it costs a host sync per step, which is why it is off by default here
        self.check_bounds = False

    # -- shared fused path ------------------------------------------------------------------------------
    def _fused_in_batch_loss(self, mode: int, q: torch.Tensor, d: torch.Tensor, offset: int) -> torch.Tensor:
        bounds = torch.empty(2, dtype=torch.float32, device=q.device) if (self.check_bounds and self.normalize_scores) else None
        loss = _InBatchLossFn.apply(q, d, int(offset), mode, self.temperature, self.normalize_scores,
                                    self.pos_aware_negative_filtering, self.filter_threshold, self.filter_factor, bounds,
                                    float(self.tau) if self.use_smooth_max else 0.0)
        if bounds is not None:
            mn, mx = bounds.tolist()
            if mn < -self.norm_tol or mx > 1 + self.norm_tol:
                print(f"Scores out of bounds after normalization: min={mn:.4f}, max={mx:.4f}, tol={self.norm_tol}")
        return loss


class ColbertLoss(ColbertModule):
    """InfoNCE loss for late interaction without explicit negatives (late_interaction_losses.py:110-164)."""

    def __init__(self, temperature: float = 0.02, normalize_scores: bool = True, use_smooth_max: bool = False,
                 pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024, tau: float = 0.1,
                 norm_tol: float = 1e-3, filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, tau, norm_tol, filter_threshold, filter_factor)
        self.temperature = temperature
        self.normalize_scores = normalize_scores
        self.use_smooth_max = use_smooth_max
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        return self._fused_in_batch_loss(_lib.CPB_LOSS_CE, query_embeddings, doc_embeddings, offset)


class ColbertPairwiseCELoss(ColbertModule):
    """Pairwise softplus loss over in-batch documents (late_interaction_losses.py:255-313)."""

    def __init__(self, temperature: float = 1.0, normalize_scores: bool = True, use_smooth_max: bool = False,
                 pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024, tau: float = 0.1,
                 norm_tol: float = 1e-3, filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, tau, norm_tol, filter_threshold, filter_factor)
        self.temperature = temperature
        self.normalize_scores = normalize_scores
        self.use_smooth_max = use_smooth_max
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        return self._fused_in_batch_loss(_lib.CPB_LOSS_PAIRWISE, query_embeddings, doc_embeddings, offset)


class ColbertSigmoidLoss(ColbertModule):
    """Sigmoid loss over the in-batch score matrix (late_interaction_losses.py:401-465).  As in the reference the score
    matrix must be square and the positives sit on the diagonal (``offset`` other than 0 indexes out of range there)."""

    def __init__(self, temperature: float = 0.02, normalize_scores: bool = True, use_smooth_max: bool = False,
                 pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024, tau: float = 0.1,
                 norm_tol: float = 1e-3, filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, tau, norm_tol, filter_threshold, filter_factor)
        self.temperature = temperature
        self.normalize_scores = normalize_scores
        self.use_smooth_max = use_smooth_max
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        if offset != 0 or query_embeddings.shape[0] != doc_embeddings.shape[0]:
            raise ValueError("ColbertSigmoidLoss needs as many documents as queries and offset == 0 "
                             f"(got {query_embeddings.shape[0]} queries, {doc_embeddings.shape[0]} documents, offset {offset})")
        return self._fused_in_batch_loss(_lib.CPB_LOSS_SIGMOID, query_embeddings, doc_embeddings, offset)


class _NegativeLossBase(ColbertModule):
    _inner_mode = _lib.CPB_LOSS_CE

    def __init__(self, temperature, normalize_scores, use_smooth_max, pos_aware_negative_filtering, in_batch_term_weight,
                 max_batch_size, tau, norm_tol, filter_threshold, filter_factor):
        super().__init__(max_batch_size, tau, norm_tol, filter_threshold, filter_factor)
        self.temperature = temperature
        self.normalize_scores = normalize_scores
        self.use_smooth_max = use_smooth_max
        self.pos_aware_negative_filtering = pos_aware_negative_filtering
        self.in_batch_term_weight = in_batch_term_weight
        assert in_batch_term_weight >= 0, "in_batch_term_weight must be non-negative"
        assert in_batch_term_weight <= 1, "in_batch_term_weight must be less than 1"

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, neg_doc_embeddings: torch.Tensor,
                offset: int = 0) -> torch.Tensor:
        return _NegLossFn.apply(query_embeddings, doc_embeddings, neg_doc_embeddings, int(offset), self._inner_mode,
                                self.temperature, self.normalize_scores, self.pos_aware_negative_filtering,
                                self.filter_threshold, self.filter_factor, self.in_batch_term_weight,
                                float(self.tau) if self.use_smooth_max else 0.0)


class ColbertNegativeCELoss(_NegativeLossBase):
    """InfoNCE-style loss with explicit negative documents (late_interaction_losses.py:167-252):
    ``(1 - w) * softplus((neg - pos) / T).mean() + w * ColbertLoss``."""

    _inner_mode = _lib.CPB_LOSS_CE

    def __init__(self, temperature: float = 0.02, normalize_scores: bool = True, use_smooth_max: bool = False,
                 pos_aware_negative_filtering: bool = False, in_batch_term_weight: float = 0.5,
                 max_batch_size: int = 1024, tau: float = 0.1, norm_tol: float = 1e-3, filter_threshold: float = 0.95,
                 filter_factor: float = 0.5):
        super().__init__(temperature, normalize_scores, use_smooth_max, pos_aware_negative_filtering,
                         in_batch_term_weight, max_batch_size, tau, norm_tol, filter_threshold, filter_factor)


class ColbertPairwiseNegativeCELoss(_NegativeLossBase):
    """Pairwise loss with explicit negatives (late_interaction_losses.py:316-398):
    ``(1 - w) * softplus((neg - pos) / T).mean() + w * ColbertPairwiseCELoss``."""

    _inner_mode = _lib.CPB_LOSS_PAIRWISE

    def __init__(self, temperature: float = 0.02, normalize_scores: bool = True, use_smooth_max: bool = False,
                 pos_aware_negative_filtering: bool = False, in_batch_term_weight: float = 0.5,
                 max_batch_size: int = 1024, tau: float = 0.1, norm_tol: float = 1e-3, filter_threshold: float = 0.95,
                 filter_factor: float = 0.5):
        super().__init__(temperature, normalize_scores, use_smooth_max, pos_aware_negative_filtering,
                         in_batch_term_weight, max_batch_size, tau, norm_tol, filter_threshold, filter_factor)
