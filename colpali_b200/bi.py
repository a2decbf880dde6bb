"""Single-vector (bi-encoder) scoring, the bi-encoder losses and similarity maps on B200.

Mirrors, with the same signatures and error behaviour:

* ``BaseVisualRetrieverProcessor.score_single_vector``   colpali_engine/utils/processing_utils.py:103-130
* ``BiEncoderLoss`` :64-113, ``BiPairedEncoderLoss`` :116-168, ``BiNegativeCELoss`` :171-248, ``BiPairwiseCELoss`` :251-302,
  ``BiPairwiseNegativeCELoss`` :305-358, ``BiSigmoidLoss`` :361-418   colpali_engine/loss/bi_encoder_losses.py
* ``get_similarity_maps_from_embeddings``                colpali_engine/interpretability/similarity_map_utils.py:9-56

All three are the un-reduced / single-token case of the late-interaction contraction: a dense ``A @ B.T``.  The Bi* models
emit ONE hidden-size vector per query / page (1536+ dims, often fp32), so this runs on the fp32 CUDA-core kernel
``csrc/dense_sm100.cu`` (operands keep their dtype: nothing is demoted to bf16), and the losses reuse the in-batch loss
body of the ColBERT losses (``csrc/loss_body.cuh``) on the resulting score matrix: one launch for the scores, one for
loss + d loss / d scores, and the backward is two (three with explicit negatives) more dense products over strided views
-- no transposed copies.  The loss is returned in fp32 whatever the embedding dtype (the reference returns the embedding
dtype).
"""

from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple, Union

import torch

from . import _lib
from .scoring import _on_device, _require_cuda, _resolve_device

TensorOrList = Union[torch.Tensor, List[torch.Tensor]]
_DOT_DTYPES = (torch.bfloat16, torch.float32)


def _as_dot_operand(x: torch.Tensor) -> torch.Tensor:
    """bf16 and fp32 go to the kernel as they are; other float types (fp16, fp64) are widened / narrowed to fp32."""
    return x if x.dtype in _DOT_DTYPES else x.to(torch.float32)


def dense_dot(a: torch.Tensor, b: torch.Tensor, *, b_rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              alpha: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """``out[i, j] (+)= alpha * sum_k a[i, k] * b[row(j), k]`` in fp32 (``cpb_dense_dot_launch``).

    ``a`` ``[m, k]`` and ``b`` ``[n_b, k]`` are 2-D CUDA tensors of dtype bf16 or fp32 with ARBITRARY strides (pass ``x.t()``
    for a transposed operand: nothing is copied); ``b_rows`` (int32 ``[n]``) gathers rows of ``b``; ``alpha`` is a device
    fp32 scalar.  Returns fp32 ``[m, n]``."""
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise ValueError(f"dense_dot needs [m, k] and [n, k] operands, got {tuple(a.shape)} and {tuple(b.shape)}")
    dev = a.device
    _require_cuda(dev)
    if b.device != dev:
        raise _lib.ColpaliB200Error("dense_dot operands live on different devices")
    a, b = _as_dot_operand(a), _as_dot_operand(b)
    m, k = a.shape
    if b_rows is not None and (b_rows.device != dev or b_rows.dtype != torch.int32 or not b_rows.is_contiguous()):
        raise ValueError("b_rows must be a contiguous int32 tensor on the operands' device")
    if alpha is not None and (alpha.device != dev or alpha.dtype != torch.float32 or alpha.numel() != 1):
        raise ValueError("alpha must be a one-element fp32 tensor on the operands' device")
    if out is not None and out.device != dev:
        raise ValueError("the output tensor lives on another device")
    n = int(b_rows.numel()) if b_rows is not None else b.shape[0]
    if out is None:
        if accumulate:
            raise ValueError("accumulate=True needs an output tensor")
        out = torch.empty(m, n, dtype=torch.float32, device=dev)
    if out.dtype != torch.float32 or out.shape != (m, n) or out.stride(1) != 1:
        raise ValueError("dense_dot output must be an fp32 [m, n] tensor with unit column stride")
    if m == 0 or n == 0:
        return out
    if k == 0:
        return out if accumulate else out.zero_()
    args = _lib.DenseDotArgs()
    args.flags = ((_lib.CPB_DOT_A_F32 if a.dtype == torch.float32 else 0) | (_lib.CPB_DOT_B_F32 if b.dtype == torch.float32 else 0)
                  | (_lib.CPB_DOT_ACCUMULATE if accumulate else 0))
    args.d_a, args.a_row_stride, args.a_k_stride = a.data_ptr(), a.stride(0), a.stride(1)
    args.d_b, args.b_row_stride, args.b_k_stride = b.data_ptr(), b.stride(0), b.stride(1)
    args.d_b_rows = b_rows.data_ptr() if b_rows is not None else None
    args.m, args.n, args.k = m, n, k
    args.d_out, args.out_row_stride = out.data_ptr(), out.stride(0)
    args.d_alpha = alpha.data_ptr() if alpha is not None else None
    with _on_device(dev):
        args.stream = torch.cuda.current_stream(dev).cuda_stream
        rc = _lib.load().cpb_dense_dot_launch(ctypes.byref(args))
    _lib.check(rc, "cpb_dense_dot_launch")
    _lib.count_launches(1)
    return out


def score_single_vector(qs: TensorOrList, ps: TensorOrList,
                        device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
    """Drop-in for ``BaseVisualRetrieverProcessor.score_single_vector`` (processing_utils.py:103-130): fp32
    ``[n_queries, n_passages]`` dot products ON THE DEVICE (the reference does not move them to the CPU either), same
    ``ValueError``s on empty lists.  Any embedding dim, bf16 or fp32 operands at full precision."""
    dev = _resolve_device(device)
    _require_cuda(dev)
    if isinstance(qs, list) and isinstance(ps, list):
        if len(qs) == 0:
            raise ValueError("No queries provided")
        if len(ps) == 0:
            raise ValueError("No passages provided")
        qs, ps = torch.stack(qs), torch.stack(ps)
    qs, ps = qs.to(dev), ps.to(dev)
    if qs.dim() != 2 or ps.dim() != 2:
        raise ValueError(f"single-vector embeddings must be [n, dim], got {tuple(qs.shape)} and {tuple(ps.shape)}")
    scores = dense_dot(qs, ps)
    assert scores.shape[0] == len(qs), f"Expected {len(qs)} scores, got {scores.shape[0]}"
    return scores


def get_similarity_maps_from_embeddings(image_embeddings: torch.Tensor, query_embeddings: torch.Tensor,
                                        n_patches: Union[Tuple[int, int], List[Tuple[int, int]]],
                                        image_mask: torch.Tensor) -> List[torch.Tensor]:
    """Drop-in for similarity_map_utils.py:9-56: one ``(query_tokens, n_patches_x, n_patches_y)`` fp32 map per image, the
    un-reduced late-interaction products ``einsum("nk,ijk->nij")`` (:50-52).  The mask selection (:43) and the
    ``"(h w) c -> w h c"`` regrouping (:42-47) become the row-gather index of one dense product per image: the patch grid is
    never materialised."""
    if isinstance(n_patches, tuple):
        n_patches = [n_patches] * image_embeddings.size(0)
    dev = image_embeddings.device
    _require_cuda(dev)
    maps: List[torch.Tensor] = []
    counts = image_mask.sum(dim=1).tolist()  # one host read for the sanity check of :34-40
    for idx in range(image_embeddings.size(0)):
        w, h = int(n_patches[idx][0]), int(n_patches[idx][1])
        if counts[idx] != w * h:
            raise ValueError(
                f"The number of patches ({w} x {h} = {w * h}) "
                f"does not match the number of non-padded image tokens ({counts[idx]})."
            )
        rows = torch.nonzero(image_mask[idx], as_tuple=False).flatten()      # masked tokens, row-major "(h w)"
        rows = rows.view(h, w).t().contiguous().view(-1).to(device=dev, dtype=torch.int32)  # output order: i = w index, j = h index
        sim = dense_dot(query_embeddings[idx].to(dev), image_embeddings[idx], b_rows=rows)
        maps.append(sim.view(query_embeddings.shape[1], w, h))
    return maps


# ------------------------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------------------------
def _check_vectors(q, d, neg=None):
    if q.dim() != 2 or d.dim() != 2 or (neg is not None and neg.dim() != 3):
        raise ValueError("expected [B, D], [C, D] (and [B, n_neg, D]) embeddings, got "
                         f"{tuple(q.shape)} / {tuple(d.shape)}" + (f" / {tuple(neg.shape)}" if neg is not None else ""))
    dev = q.device
    if dev.type != "cuda" or d.device != dev or (neg is not None and neg.device != dev):
        raise _lib.ColpaliB200Error("colpali_b200 losses need all embeddings on the same CUDA device")
    return dev


class _BiLossFn(torch.autograd.Function):
    """scores = q d^T (and q against every query's negatives) -> loss body -> loss and d loss / d scores; backward =
    dense products of those gradients with the saved embeddings."""

    @staticmethod
    def forward(ctx, q, d, neg, offset, neg_delta, mode, temperature, filt, thr, factor, weight):
        dev = _check_vectors(q, d, neg)
        qd, dd = q.detach(), d.detach()
        b, c = qd.shape[0], dd.shape[0]
        need_grad = any(ctx.needs_input_grad[:3])
        scores = dense_dot(qd, dd)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g = torch.empty(b, c, dtype=torch.float32, device=dev) if need_grad else None
        desc = _lib.LossDesc()
        desc.mode, desc.normalize_scores, desc.pos_aware_negative_filtering, desc.offset = int(mode), 0, int(bool(filt)), int(offset)
        desc.temperature, desc.filter_threshold, desc.filter_factor = float(temperature), float(thr), float(factor)
        desc.d_loss = loss.data_ptr()
        desc.d_grad_scores = g.data_ptr() if g is not None else None
        nflat = s_neg = g_neg = None
        if neg is not None:
            if neg.shape[0] != b:
                raise ValueError(f"{neg.shape[0]} negative groups for {b} queries")
            n_neg = neg.shape[1]
            nflat = neg.detach().reshape(b * n_neg, neg.shape[2])
            s_neg = dense_dot(qd, nflat)  # every query against every query's negatives; the loss reads the block diagonal
            g_neg = torch.empty(b, b * n_neg, dtype=torch.float32, device=dev) if need_grad else None
            desc.d_neg_scores, desc.n_neg, desc.in_batch_term_weight = s_neg.data_ptr(), int(n_neg), float(weight)
            desc.d_grad_neg_scores = g_neg.data_ptr() if g_neg is not None else None
            desc.neg_pos_offset_delta = int(neg_delta)
        with _on_device(dev):
            rc = _lib.load().cpb_colbert_loss_launch(ctypes.byref(desc), scores.data_ptr(), None, b, 0, c, 0,
                                                     torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "cpb_colbert_loss_launch")
        _lib.count_launches(1)
        if need_grad:
            ctx.save_for_backward(qd, dd, nflat, g, g_neg)
            ctx.neg_shape = tuple(neg.shape) if neg is not None else None
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        q, d, nflat, g, g_neg = ctx.saved_tensors
        want_q, want_d, want_n = ctx.needs_input_grad[:3]
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        grad_q = grad_d = grad_n = None
        if want_q:   # dQ = G D (+ Gn N)
            dq = dense_dot(g, d.t(), alpha=go)
            if nflat is not None:
                dense_dot(g_neg, nflat.t(), alpha=go, out=dq, accumulate=True)
            grad_q = dq.to(q.dtype)
        if want_d:   # dD = G^T Q
            grad_d = dense_dot(g.t(), q.t(), alpha=go).to(d.dtype)
        if want_n and nflat is not None:  # dN = Gn^T Q (zero outside each query's own negatives)
            grad_n = dense_dot(g_neg.t(), q.t(), alpha=go).to(nflat.dtype).view(ctx.neg_shape)
        return (grad_q, grad_d, grad_n) + (None,) * 8


class BiEncoderModule(torch.nn.Module):
    """bi_encoder_losses.py:6-61 -- hyper-parameters shared by the bi-encoder losses (the eager helpers ``_get_idx`` /
    ``_filter_high_negatives`` live inside the loss kernel here)."""

    def __init__(self, max_batch_size: int = 1024, temperature: float = 0.02, filter_threshold: float = 0.95,
                 filter_factor: float = 0.5):
        super().__init__()
        if temperature <= 0:
            raise ValueError("Temperature must be strictly positive")  # :25-26
        self.register_buffer("idx_buffer", torch.arange(max_batch_size), persistent=False)
        self.temperature = temperature
        self.filter_threshold = filter_threshold
        self.filter_factor = filter_factor

    def _launch(self, mode, q, d, neg=None, offset=0, neg_delta=0, filt=False, weight=1.0):
        return _BiLossFn.apply(q, d, neg, int(offset), int(neg_delta), mode, self.temperature, filt, self.filter_threshold,
                               self.filter_factor, weight)


class BiEncoderLoss(BiEncoderModule):
    """InfoNCE over in-batch documents (bi_encoder_losses.py:64-113)."""

    def __init__(self, temperature: float = 0.02, pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, temperature, filter_threshold, filter_factor)
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        return self._launch(_lib.CPB_LOSS_CE, query_embeddings, doc_embeddings, offset=offset,
                            filt=self.pos_aware_negative_filtering)


class BiPairedEncoderLoss(BiEncoderModule):
    """Symmetric InfoNCE, (query->doc + doc->query) / 2 (bi_encoder_losses.py:116-168).  As in the reference the score
    matrix must be square with the positives on the diagonal (its ``CrossEntropyLoss(scores.T, idx)`` fails otherwise)."""

    def __init__(self, temperature: float = 0.02, pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, temperature, filter_threshold, filter_factor)
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        if offset != 0 or query_embeddings.shape[0] != doc_embeddings.shape[0]:
            raise ValueError("BiPairedEncoderLoss needs as many documents as queries and offset == 0 "
                             f"(got {query_embeddings.shape[0]} queries, {doc_embeddings.shape[0]} documents, offset {offset})")
        return self._launch(_lib.CPB_LOSS_SYMMETRIC_CE, query_embeddings, doc_embeddings,
                            filt=self.pos_aware_negative_filtering)


class BiNegativeCELoss(BiEncoderModule):
    """``(1 - w) * softplus((neg - pos) / T).mean() + w * BiEncoderLoss`` (bi_encoder_losses.py:171-248)."""

    def __init__(self, temperature: float = 0.02, in_batch_term_weight: float = 0.5,
                 pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024, filter_threshold: float = 0.95,
                 filter_factor: float = 0.5):
        super().__init__(max_batch_size, temperature, filter_threshold, filter_factor)
        self.in_batch_term_weight = in_batch_term_weight
        assert 0 <= in_batch_term_weight <= 1, "in_batch_term_weight must be between 0 and 1"
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, neg_doc_embeddings: torch.Tensor,
                offset: int = 0) -> torch.Tensor:
        return self._launch(_lib.CPB_LOSS_CE, query_embeddings, doc_embeddings, neg_doc_embeddings, offset=offset,
                            filt=self.pos_aware_negative_filtering, weight=self.in_batch_term_weight)


class BiPairwiseCELoss(BiEncoderModule):
    """softplus((hardest in-batch negative - positive) / T) (bi_encoder_losses.py:251-302).  Like the reference, the
    positives are ``scores.diagonal()``: ``offset`` is accepted and IGNORED (:283-292)."""

    def __init__(self, temperature: float = 0.02, pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, temperature, filter_threshold, filter_factor)
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        return self._launch(_lib.CPB_LOSS_PAIRWISE, query_embeddings, doc_embeddings, offset=0,
                            filt=self.pos_aware_negative_filtering)


class BiPairwiseNegativeCELoss(BiEncoderModule):
    """``(1 - w) * softplus((neg - pos) / T).mean() + w * BiPairwiseCELoss`` (bi_encoder_losses.py:305-358).  The explicit
    term takes its positives at ``offset`` (:350); the in-batch term, being BiPairwiseCELoss, ignores it (:355)."""

    def __init__(self, temperature: float = 0.02, in_batch_term_weight: float = 0.5, max_batch_size: int = 1024,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, temperature, filter_threshold, filter_factor)
        self.in_batch_term_weight = in_batch_term_weight
        assert 0 <= in_batch_term_weight <= 1, "in_batch_term_weight must be between 0 and 1"

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, neg_doc_embeddings: torch.Tensor,
                offset: int = 0) -> torch.Tensor:
        return self._launch(_lib.CPB_LOSS_PAIRWISE, query_embeddings, doc_embeddings, neg_doc_embeddings, offset=0,
                            neg_delta=offset, filt=False, weight=self.in_batch_term_weight)


class BiSigmoidLoss(BiEncoderModule):
    """Sigmoid loss over all (query, document) pairs (bi_encoder_losses.py:361-418): +1 at the positive (column b + offset),
    -1 elsewhere, mean over the ``B x C`` matrix.  The reference walks ``C // B`` column blocks starting at ``offset``; that
    visits every column exactly once when ``C`` and ``offset`` are multiples of ``B`` (and indexes out of range or drops
    columns otherwise), which is what is required here."""

    def __init__(self, temperature: float = 0.02, pos_aware_negative_filtering: bool = False, max_batch_size: int = 1024,
                 filter_threshold: float = 0.95, filter_factor: float = 0.5):
        super().__init__(max_batch_size, temperature, filter_threshold, filter_factor)
        self.pos_aware_negative_filtering = pos_aware_negative_filtering

    def forward(self, query_embeddings: torch.Tensor, doc_embeddings: torch.Tensor, offset: int = 0) -> torch.Tensor:
        b, c = query_embeddings.shape[0], doc_embeddings.shape[0]
        if b == 0 or c % b != 0 or offset % b != 0:
            raise ValueError(f"BiSigmoidLoss needs n_docs ({c}) and offset ({offset}) to be multiples of n_queries ({b})")
        return self._launch(_lib.CPB_LOSS_SIGMOID, query_embeddings, doc_embeddings, offset=offset,
                            filt=self.pos_aware_negative_filtering)
