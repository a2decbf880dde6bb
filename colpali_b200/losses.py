"""ColBERT in-batch-negative losses on B200 behind the reference's loss-module API.

Mirrors ``colpali_engine/loss/late_interaction_losses.py`` (reference @ 9be8f19):

* ``ColbertModule``           :6-107   (hyper-parameters only; the reference's eager helper methods ``_aggregate`` /
                                        ``_smooth_max`` / ``_filter_high_negatives`` live inside the kernels here)
* ``ColbertLoss``             :110-164 (InfoNCE over in-batch documents)
* ``ColbertPairwiseCELoss``   :255-313 (softplus(hardest in-batch negative - positive))
* ``ColbertNegativeCELoss``   :167-252, ``ColbertPairwiseNegativeCELoss`` :316-398 (explicit negatives)
* ``ColbertSigmoidLoss``      :401-465

Same constructor keyword arguments, same ``forward(query_embeddings, doc_embeddings, offset=0)``.  The
``einsum("bnd,csd->bcns")`` / ``amax`` / ``sum`` of :153-154 and everything after it run as

1. the fused sm_100a MaxSim kernel (csrc/maxsim_sm100.cu), which also records, per (document, query token),
   the index of the winning document token (int32 ``[C, B*N_q]`` instead of the reference's saved
   ``[B, C, N_q, N_d]`` similarity tensor);
2. one small kernel (csrc/loss_sm100.cu) that turns the ``[B, C]`` sums into the scalar loss *and* its gradient
   with respect to the sums;
3. in backward, a gather kernel for ``dQ`` and a scatter-add kernel for ``dD``.

Differences from the reference that a caller can observe, all documented in DESIGN.md:
the loss is returned in fp32 whatever the embedding dtype (the reference returns the embedding dtype);
embeddings are contracted in bf16 with fp32 accumulation; ``use_smooth_max=True`` is not implemented yet and
raises; where several document tokens tie for the maximum the gradient goes to the first one (``amax`` splits
it evenly) -- this only differs on all-zero (padding) rows, whose gradients the model masks anyway.
"""

from __future__ import annotations

import torch

from . import _lib
from .scoring import DocBank, QueryBlock, maxsim


class _InBatchLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, d, offset, mode, temperature, normalize, filt, thr, factor, bounds_out):
        if q.dim() != 3 or d.dim() != 3:
            raise ValueError(f"expected [B, N_q, D] and [C, N_d, D] embeddings, got {tuple(q.shape)} / {tuple(d.shape)}")
        dev = q.device
        if dev.type != "cuda" or d.device != dev:
            raise _lib.ColpaliB200Error("colpali_b200 losses need query and document embeddings on the same CUDA device")
        lib = _lib.load()
        qb = QueryBlock(q.detach(), dev)
        bank = DocBank.from_passages(d.detach(), dev)  # dense [C, L, D]: zero (padding) rows are ordinary tokens,
        # they score exactly 0 and take part in the max, as in the reference
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        if need_grad:
            scores, argmax = maxsim(qb, bank, want_argmax=True)
        else:
            scores, argmax = maxsim(qb, bank), None
        b, c = qb.n, bank.n_docs
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g = torch.empty(b, c, dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            rc = lib.cpb_colbert_loss_fwd_dim(
                scores.data_ptr(), qb.flat.data_ptr(), b, qb.nq_pad, c, mode,
                float(temperature), int(normalize), int(filt), float(thr), float(factor), int(offset),
                loss.data_ptr(), g.data_ptr() if g is not None else None,
                bounds_out.data_ptr() if bounds_out is not None else None,
                qb.flat.shape[1], torch.cuda.current_stream(dev).cuda_stream,
            )
        _lib.check(rc, "cpb_colbert_loss_fwd_dim")
        _lib.count_launches(1)
        if need_grad:
            ctx.save_for_backward(qb.flat, bank.flat, bank.start, argmax, g)
            ctx.meta = (qb.n, qb.nq_pad, c, tuple(q.shape), tuple(d.shape), q.dtype, d.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        q_flat, d_flat, d_start, argmax, g = ctx.saved_tensors
        b, nq_pad, c, q_shape, d_shape, q_dtype, d_dtype = ctx.meta
        dev = q_flat.device
        lib = _lib.load()
        want_q, want_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dim = q_flat.shape[1]  # padded embedding dim (128, or 192 / 256 / 320 for wide models)
        dq = torch.empty(b * nq_pad, dim, dtype=torch.float32, device=dev) if want_q else None
        dd = torch.zeros(d_flat.shape[0], dim, dtype=torch.float32, device=dev) if want_d else None
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            rc = lib.cpb_maxsim_bwd_dim(
                g.data_ptr(), go.data_ptr(), argmax.data_ptr(),
                q_flat.data_ptr(), b, nq_pad,
                d_flat.data_ptr(), d_flat.shape[0], d_start.data_ptr(), c,
                dq.data_ptr() if dq is not None else None, dd.data_ptr() if dd is not None else None,
                dim, torch.cuda.current_stream(dev).cuda_stream,
            )
        _lib.check(rc, "cpb_maxsim_bwd_dim")
        _lib.count_launches(int(want_q) + int(want_d))
        grad_q = grad_d = None
        if want_q:
            grad_q = dq.view(b, nq_pad, dim)[:, : q_shape[1], : q_shape[2]].to(q_dtype)
        if want_d:
            grad_d = dd.view(d_shape[0], d_shape[1], dim)[:, :, : d_shape[2]].to(d_dtype)
        return grad_q, grad_d, None, None, None, None, None, None, None, None


class _NegLossFn(torch.autograd.Function):
    """Explicit-negative losses: one dense MaxSim launch against the (gathered) positives, one against the flattened
    negatives of all queries (only the block diagonal is used), one loss kernel for both score matrices."""

    @staticmethod
    def forward(ctx, q, d, neg, offset, inner_mode, temperature, normalize, filt, thr, factor, weight):
        if q.dim() != 3 or d.dim() != 3 or neg.dim() != 4:
            raise ValueError("expected [B, N_q, D], [C, N_d, D] and [B, n_neg, N_neg, D] embeddings, got "
                             f"{tuple(q.shape)} / {tuple(d.shape)} / {tuple(neg.shape)}")
        if neg.shape[0] != q.shape[0]:
            raise ValueError(f"{neg.shape[0]} negative groups for {q.shape[0]} queries")
        dev = q.device
        if dev.type != "cuda" or d.device != dev or neg.device != dev:
            raise _lib.ColpaliB200Error("colpali_b200 losses need all embeddings on the same CUDA device")
        lib = _lib.load()
        qb = QueryBlock(q.detach(), dev)
        bank = DocBank.from_passages(d.detach(), dev)
        b, n_neg = neg.shape[0], neg.shape[1]
        nbank = DocBank.from_passages(neg.detach().reshape(b * n_neg, neg.shape[2], neg.shape[3]), dev)
        need_grad = any(ctx.needs_input_grad[:3])
        if need_grad:
            s_pos, am_pos = maxsim(qb, bank, want_argmax=True)
            s_neg, am_neg = maxsim(qb, nbank, want_argmax=True)
        else:
            s_pos, s_neg, am_pos, am_neg = maxsim(qb, bank), maxsim(qb, nbank), None, None
        c = bank.n_docs
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g_pos = torch.empty(b, c, dtype=torch.float32, device=dev) if need_grad else None
        g_neg = torch.empty(b, b * n_neg, dtype=torch.float32, device=dev) if need_grad else None
        with torch.cuda.device(dev):
            rc = lib.cpb_colbert_neg_loss_fwd_dim(
                s_pos.data_ptr(), s_neg.data_ptr(), qb.flat.data_ptr(), b, qb.nq_pad, c, n_neg, inner_mode,
                float(temperature), int(normalize), int(filt), float(thr), float(factor), float(weight), int(offset),
                loss.data_ptr(), g_pos.data_ptr() if need_grad else None, g_neg.data_ptr() if need_grad else None,
                qb.flat.shape[1], torch.cuda.current_stream(dev).cuda_stream,
            )
        _lib.check(rc, "cpb_colbert_neg_loss_fwd_dim")
        _lib.count_launches(1)
        if need_grad:
            ctx.save_for_backward(qb.flat, bank.flat, bank.start, am_pos, g_pos, nbank.flat, nbank.start, am_neg, g_neg)
            ctx.meta = (b, qb.nq_pad, c, b * n_neg, tuple(q.shape), tuple(d.shape), tuple(neg.shape), q.dtype, d.dtype, neg.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        q_flat, d_flat, d_start, am_pos, g_pos, n_flat, n_start, am_neg, g_neg = ctx.saved_tensors
        b, nq_pad, c, cn, q_shape, d_shape, n_shape, q_dtype, d_dtype, n_dtype = ctx.meta
        dev = q_flat.device
        lib = _lib.load()
        want_q, want_d, want_n = ctx.needs_input_grad[:3]
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        dim = q_flat.shape[1]

        def bwd(g, am, flat, start, n_docs, need_dq, need_dd):
            dq = torch.empty(b * nq_pad, dim, dtype=torch.float32, device=dev) if need_dq else None
            dd = torch.zeros(flat.shape[0], dim, dtype=torch.float32, device=dev) if need_dd else None
            if need_dq or need_dd:
                with torch.cuda.device(dev):
                    rc = lib.cpb_maxsim_bwd_dim(g.data_ptr(), go.data_ptr(), am.data_ptr(), q_flat.data_ptr(), b, nq_pad,
                                                flat.data_ptr(), flat.shape[0], start.data_ptr(), n_docs,
                                                dq.data_ptr() if need_dq else None, dd.data_ptr() if need_dd else None,
                                                dim, stream)
                _lib.check(rc, "cpb_maxsim_bwd_dim")
                _lib.count_launches(int(need_dq) + int(need_dd))
            return dq, dd

        dq1, dd = bwd(g_pos, am_pos, d_flat, d_start, c, want_q, want_d)
        dq2, dn = bwd(g_neg, am_neg, n_flat, n_start, cn, want_q, want_n)
        grad_q = grad_d = grad_n = None
        if want_q:
            grad_q = (dq1 + dq2).view(b, nq_pad, dim)[:, : q_shape[1], : q_shape[2]].to(q_dtype)
        if want_d:
            grad_d = dd.view(d_shape[0], d_shape[1], dim)[:, :, : d_shape[2]].to(d_dtype)
        if want_n:
            grad_n = dn.view(n_shape[0], n_shape[1], n_shape[2], dim)[..., : n_shape[3]].to(n_dtype)
        return (grad_q, grad_d, grad_n) + (None,) * 8


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
        if self.use_smooth_max:
            raise NotImplementedError(
                "use_smooth_max=True (tau * logsumexp instead of amax, late_interaction_losses.py:40-44) is not "
                "implemented in the fused sm_100a path yet; use the reference module for it."
            )
        bounds = torch.empty(2, dtype=torch.float32, device=q.device) if (self.check_bounds and self.normalize_scores) else None
        loss = _InBatchLossFn.apply(q, d, int(offset), mode, self.temperature, self.normalize_scores,
                                    self.pos_aware_negative_filtering, self.filter_threshold, self.filter_factor, bounds)
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
        if self.use_smooth_max:
            raise NotImplementedError("use_smooth_max=True is not implemented in the fused sm_100a path yet")
        return _NegLossFn.apply(query_embeddings, doc_embeddings, neg_doc_embeddings, int(offset), self._inner_mode,
                                self.temperature, self.normalize_scores, self.pos_aware_negative_filtering,
                                self.filter_threshold, self.filter_factor, self.in_batch_term_weight)


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
