"""Late-interaction (MaxSim) scoring on B200 behind the reference's scorer API.

Mirrors ``BaseVisualRetrieverProcessor.score_multi_vector``
(/root/reference/colpali_engine/utils/processing_utils.py:132-187): same signature, same
``ValueError``s, returns a CPU fp32 ``[n_queries, n_passages]`` tensor.  The arithmetic of
line 179 (``einsum("bnd,csd->bcns").max(dim=3)[0].sum(dim=2)``) runs in ONE fused sm_100a kernel
(csrc/maxsim_sm100.cu) over a flat, device-resident document bank; the reference's re-padding and
re-upload of every 128-document batch for every 128-query batch (:170-180) disappears.

Zero-padding semantics are kept: the reference pads each 128-document batch with zero rows up to
that batch's longest document, and those rows take part in the max (a shorter document's
per-token maximum is ``max(real, 0)``).  ``DocBank`` records that as a per-document floor.
"""

from __future__ import annotations

import ctypes
from typing import List, Optional, Union

import torch

from . import _lib

EMBED_DIM = 128  # embedding dim one kernel pass contracts over (smaller dims are zero-padded)
TensorOrList = Union[torch.Tensor, List[torch.Tensor]]


def _require_cuda(device: torch.device) -> None:
    if device.type != "cuda":
        raise _lib.ColpaliB200Error(
            f"colpali_b200 computes on CUDA sm_100a only (got device '{device}'); there is no CPU path. "
            "Use the reference implementation for CPU scoring."
        )


def _resolve_device(device: Optional[Union[str, torch.device]]) -> torch.device:
    # processing_utils.py:161 / utils/torch_utils.py:22-29: None -> "cuda:0" when available.
    if device is None or (isinstance(device, str) and device == "auto"):
        if not torch.cuda.is_available():
            raise _lib.ColpaliB200Error("no CUDA device available; colpali_b200 has no CPU path")
        return torch.device("cuda:0")
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


MAX_EMBED_DIM = 320  # ColQwen3; dims above 128 go through the K-pipelined kernel (csrc/maxsim_kpipe_sm100.cu)


def _padded_dim(d: int) -> int:
    if d <= EMBED_DIM:
        return EMBED_DIM
    if d > MAX_EMBED_DIM:
        raise _lib.ColpaliB200Error(f"embedding dim {d} > {MAX_EMBED_DIM} is not supported by this build")
    return (d + 63) // 64 * 64


def _pad_dim(x: torch.Tensor) -> torch.Tensor:
    """bf16, last dim zero-padded to 128 (or to the next multiple of 64 above 128): zero columns add nothing to a dot product."""
    d = x.shape[-1]
    target = _padded_dim(d)
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    if d < target:
        x = torch.nn.functional.pad(x, (0, target - d))
    return x


class QueryBlock:
    """Queries laid out for the kernel: ``[n * nq_pad, 128]`` bf16, each query padded with zero
    rows to ``nq_pad`` (multiple of 32) rows, so that one epilogue warp == one query segment."""

    def __init__(self, qs: TensorOrList, device: torch.device):
        _require_cuda(device)
        if isinstance(qs, torch.Tensor):
            if qs.dim() != 3:
                raise ValueError(f"query tensor must be [n, len, dim], got {tuple(qs.shape)}")
            n, nq, _ = qs.shape
            lens = [nq] * n
        else:
            n = len(qs)
            lens = [int(q.shape[0]) for q in qs]
            nq = max(lens) if n else 0
        if n == 0:
            raise ValueError("No queries provided")
        self.n = n
        self.nq_pad = max(32, (nq + 31) // 32 * 32)
        self.device = device
        raw_dim = qs.shape[2] if isinstance(qs, torch.Tensor) else int(qs[0].shape[-1])
        self.dim = _padded_dim(raw_dim)
        if (isinstance(qs, torch.Tensor) and qs.dtype == torch.bfloat16 and qs.is_contiguous()
                and nq == self.nq_pad and qs.shape[2] == self.dim):
            if qs.device == device and qs.data_ptr() % 16 == 0:
                self.flat = qs.view(n * nq, self.dim)  # already in kernel layout: zero copy
                return
            if qs.device != device:  # kernel layout on another device (the host): one upload, no zero-fill / re-pad pass
                self.flat = qs.to(device, non_blocking=True).view(n * nq, self.dim)
                return
        flat = torch.zeros(n, self.nq_pad, self.dim, dtype=torch.bfloat16, device=device)
        if isinstance(qs, torch.Tensor):
            flat[:, :nq] = _pad_dim(qs.to(device, non_blocking=True))
        else:
            for i, q in enumerate(qs):
                if lens[i]:
                    flat[i, : lens[i]] = _pad_dim(q.to(device, non_blocking=True))
        self.flat = flat.view(n * self.nq_pad, self.dim)


_DENSE_LAYOUT_CACHE: dict = {}


def _dense_layout(n: int, length: int, device: torch.device):
    """(start, len) int32 device arrays of a dense [n, length, dim] bank; cached, they only depend on the shape."""
    key = (n, length, str(device))
    hit = _DENSE_LAYOUT_CACHE.get(key)
    if hit is None:
        if len(_DENSE_LAYOUT_CACHE) > 64:
            _DENSE_LAYOUT_CACHE.clear()
        hit = (torch.arange(0, n * length, length, dtype=torch.int32, device=device),
               torch.full((n,), length, dtype=torch.int32, device=device))
        _DENSE_LAYOUT_CACHE[key] = hit
    return hit


class DocBank:
    """Device-resident document bank: flat ``[rows, 128]`` bf16 tokens + (start, len, floor) per doc.

    ``floor[j]`` is 0 where the reference would have zero-padded document ``j`` inside its
    128-document batch (processing_utils.py:176-178) and -inf elsewhere.  Build once, score many
    query batches against it.
    """

    def __init__(self, flat: torch.Tensor, start: torch.Tensor, length: torch.Tensor,
                 floor: Optional[torch.Tensor], contiguous: bool = False, uniform_len: int = 0, max_len: int = 0):
        self.flat, self.start, self.length, self.floor = flat, start, length, floor
        self.contiguous = contiguous  # documents stored back to back (start[j+1] == start[j] + len[j])
        self.uniform_len = uniform_len  # > 0: every document has this many rows
        self.max_len = max_len          # longest document (0 = unknown: no tile-balanced partitioning)
        self.n_docs = int(start.numel())
        self.device = flat.device

    def __len__(self) -> int:
        return self.n_docs

    @staticmethod
    def from_passages(ps: TensorOrList, device: torch.device, batch_size: int = 128,
                      reference_padding: bool = True) -> "DocBank":
        _require_cuda(device)
        if len(ps) == 0:
            raise ValueError("No passages provided")
        if isinstance(ps, torch.Tensor):
            if ps.dim() != 3:
                raise ValueError(f"passage tensor must be [n, len, dim], got {tuple(ps.shape)}")
            n, L, _ = ps.shape
            flat = _pad_dim(ps.to(device, non_blocking=True))
            flat = flat.reshape(n * L, flat.shape[-1]).contiguous()
            start, length = _dense_layout(n, L, device)
            return DocBank(flat, start, length, None, contiguous=True, uniform_len=L, max_len=L)  # equal lengths: no padding
        lens = [int(p.shape[0]) for p in ps]
        n = len(ps)
        # one pass over the bank: device-side cat of the per-document uploads
        flat = _pad_dim(torch.cat([p.to(device, non_blocking=True) for p in ps], dim=0)).contiguous()
        if flat.shape[0] == 0:
            flat = torch.zeros(1, flat.shape[-1], dtype=torch.bfloat16, device=device)
        lens_t = torch.tensor(lens, dtype=torch.int64)
        start = (torch.cumsum(lens_t, 0) - lens_t).to(torch.int32).to(device)
        length = lens_t.to(torch.int32).to(device)
        floor = None
        if reference_padding:
            fl = torch.full((n,), float("-inf"), dtype=torch.float32)
            any_pad = False
            for j in range(0, n, batch_size):
                chunk = lens_t[j : j + batch_size]
                padded = chunk < chunk.max()
                if bool(padded.any()):
                    any_pad = True
                    fl[j : j + batch_size][padded] = 0.0
            floor = fl.to(device) if any_pad else None
        uniform = lens[0] if len(set(lens)) == 1 else 0
        return DocBank(flat, start, length, floor, contiguous=True, uniform_len=uniform, max_len=max(lens))


_SPLIT_WS: dict = {}
_EPOCH = [0]


def _split_workspace(dev: torch.device, stream_id: int, nbytes: int) -> torch.Tensor:
    """Exchange buffer for documents cut by a partition boundary, one per (device, stream): launches that share it
    are stream-ordered (zeroed once, reused: slots are tagged with a per-launch epoch)."""
    key = (str(dev), stream_id)
    ws = _SPLIT_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        _SPLIT_WS[key] = ws
    return ws


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL_CTX = _NullCtx()


def _on_device(dev: torch.device):
    """``torch.cuda.device(dev)`` only when it is not already the current device (the context manager costs ~8 us of
    host time per launch, comparable to the small kernels of the loss path)."""
    return _NULL_CTX if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


def next_epoch() -> int:
    _EPOCH[0] = _EPOCH[0] % 0xFFFFFFF0 + 1
    return _EPOCH[0]


def launch_maxsim(q: "QueryBlock", bank: "DocBank", *, scores: Optional[torch.Tensor], argmax: Optional[torch.Tensor] = None,
                  lse: Optional[torch.Tensor] = None, round_bf16: bool = False, independent: bool = False,
                  smooth_tau: float = 0.0, nq_real: int = 0, loss: Optional[_lib.LossDesc] = None,
                  done_counter: Optional[torch.Tensor] = None, gather: Optional[dict] = None,
                  wait: Optional[tuple] = None, topk: Optional[tuple] = None) -> int:
    """Fill a ``cpb_maxsim_args`` and enqueue the fused kernel on the current stream of ``bank.device``.
    Returns the number of CTAs launched (the fused all-gather's consumers count completions in CTAs)."""
    lib = _lib.load()
    dev = bank.device
    dim = int(bank.flat.shape[1])
    if q.flat.shape[1] != dim:
        raise ValueError(f"queries have (padded) dim {q.flat.shape[1]}, documents {dim}")
    ws = None
    if q.nq_pad > 32:  # per-segment partial scores, summed by a second tiny kernel
        ws = torch.empty(lib.cpb_maxsim_workspace_bytes(q.n, q.nq_pad, bank.n_docs) // 4, dtype=torch.float32, device=dev)
    a = _lib.MaxSimArgs()
    a.flags = ((_lib.CPB_FLAG_ROUND_BF16 if round_bf16 else 0) | (_lib.CPB_FLAG_CONTIGUOUS if bank.contiguous else 0)
               | (_lib.CPB_FLAG_INDEPENDENT if independent else 0))
    a.d_q, a.n_queries, a.nq_pad, a.nq_real, a.dim = q.flat.data_ptr(), q.n, q.nq_pad, int(nq_real), dim
    a.d_docs, a.doc_rows = bank.flat.data_ptr(), bank.flat.shape[0]
    a.d_doc_start, a.d_doc_len = bank.start.data_ptr(), bank.length.data_ptr()
    a.d_doc_floor = bank.floor.data_ptr() if bank.floor is not None else None
    a.n_docs, a.uniform_len, a.max_doc_len = bank.n_docs, bank.uniform_len, bank.max_len
    a.d_scores = scores.data_ptr() if scores is not None else None
    a.d_argmax = argmax.data_ptr() if argmax is not None else None
    a.d_lse = lse.data_ptr() if lse is not None else None
    a.d_workspace = ws.data_ptr() if ws is not None else None
    a.smooth_tau = float(smooth_tau)
    with _on_device(dev):
        stream = torch.cuda.current_stream(dev)
        a.stream = stream.cuda_stream
        if dim == EMBED_DIM and bank.contiguous and bank.max_len > 0 and smooth_tau == 0.0:
            nbytes = lib.cpb_maxsim_split_workspace_bytes(q.n, q.nq_pad) * (4 if independent else 1)
            split = _split_workspace(dev, stream.cuda_stream, nbytes)
            a.d_split_ws, a.split_ws_bytes, a.epoch = split.data_ptr(), split.numel(), next_epoch()
        if gather is not None:
            a.d_peer_bases, a.mc_base, a.n_peers = gather["peer_bases"], gather["mc_base"], gather["n_peers"]
            a.slab_word_offset, a.flag_word_offset = gather["slab_word_offset"], gather["flag_word_offset"]
            if gather.get("wait_flags"):
                a.d_wait_flags, a.n_wait, a.wait_value = gather["wait_flags"], gather["n_peers"], gather["wait_value"]
        if wait is not None:  # (device pointer to n local counter words, n, value they must have reached)
            a.d_wait_flags, a.n_wait, a.wait_value = wait
        if loss is not None:
            a.loss = ctypes.pointer(loss)
            a.d_done_counter = done_counter.data_ptr()
        if topk is not None:  # (scores [n, k] fp32, idx [n, k] int32, workspace, k): selection fused into the kernel's tail
            a.d_topk_scores, a.d_topk_idx, a.d_topk_ws, a.topk_k = (topk[0].data_ptr(), topk[1].data_ptr(),
                                                                    topk[2].data_ptr(), int(topk[3]))
        rc = lib.cpb_maxsim_launch(ctypes.byref(a))
    _lib.check(rc, "cpb_maxsim_launch")
    _lib.count_launches(2 if ws is not None else 1)
    return int(a.grid_out)


def maxsim(q: QueryBlock, bank: DocBank, *, round_bf16: bool = False, want_argmax: bool = False,
           independent: bool = False):
    """Run the fused kernel.  Returns device fp32 ``[n_queries, n_docs]`` (and int32 argmax
    ``[n_docs, n_queries * nq_pad]`` when asked).  ``independent=True`` tells the kernel that it reads nothing the
    previous kernel on the stream wrote (next query batch against a resident bank): it may then overlap that
    kernel's tail (CPB_FLAG_INDEPENDENT)."""
    dev = bank.device
    scores = torch.empty(q.n, bank.n_docs, dtype=torch.float32, device=dev)
    argmax = torch.empty(bank.n_docs, q.n * q.nq_pad, dtype=torch.int32, device=dev) if want_argmax else None
    launch_maxsim(q, bank, scores=scores, argmax=argmax, round_bf16=round_bf16, independent=independent)
    return (scores, argmax) if want_argmax else scores


_TOPK_WS: dict = {}


def maxsim_topk(q: QueryBlock, bank: DocBank, k: int, *, round_bf16: bool = False):
    """Scores AND the per-query top-``k`` in ONE launch: the CTAs of every query-tile group select the ``k`` best
    documents of its queries from the score rows still in L2, each in its own slice, and the last one merges
    (csrc/topk_tail.cuh) -- larger score first, smaller document index on ties.  Returns ``(scores [n, n_docs] fp32, top_scores [n, k'] fp32, top_idx [n, k'] int32)`` with
    ``k' = min(k, n_docs)`` (views of the kernel's outputs: no kernel runs after the launch).  ``fused_topk_supported`` tells whether this shape can take the fused path."""
    if not fused_topk_supported(q, bank, k):
        raise _lib.ColpaliB200Error("fused top-k needs dim 128, queries of at most 32 tokens and k <= %d" % _lib.CPB_TOPK_MAX)
    dev = bank.device
    key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
    ctr = _TOPK_WS.get(key)
    if ctr is None:  # group counters + candidate lists: zero once, the kernel leaves the counters at zero
        ctr = torch.zeros(int(_lib.load().cpb_maxsim_topk_workspace_bytes()), dtype=torch.uint8, device=dev)
        _TOPK_WS[key] = ctr
    scores = torch.empty(q.n, bank.n_docs, dtype=torch.float32, device=dev)
    top_s = torch.empty(q.n, k, dtype=torch.float32, device=dev)
    top_i = torch.empty(q.n, k, dtype=torch.int32, device=dev)
    launch_maxsim(q, bank, scores=scores, round_bf16=round_bf16, topk=(top_s, top_i, ctr, k))
    kk = min(k, bank.n_docs)
    return scores, top_s[:, :kk], top_i[:, :kk]


def fused_topk_supported(q: QueryBlock, bank: DocBank, k: int) -> bool:
    return 1 <= k <= _lib.CPB_TOPK_MAX and q.nq_pad == 32 and int(bank.flat.shape[1]) == EMBED_DIM


def score_multi_vector(
    qs: TensorOrList,
    ps: TensorOrList,
    batch_size: int = 128,
    device: Optional[Union[str, torch.device]] = None,
    *,
    round_bf16: bool = False,
) -> torch.Tensor:
    """Drop-in for ``BaseVisualRetrieverProcessor.score_multi_vector`` (processing_utils.py:132-187).

    ``batch_size`` no longer bounds memory (no score tensor is materialised); it is kept because it
    defines the reference's zero-padding groups, hence the result for ragged list inputs.
    Extra keyword ``round_bf16`` reproduces the reference's bf16-rounded scores for bf16 inputs.
    ``ps`` may be a device-resident ``DocBank`` (build once, score many query batches).

    Numerics: embeddings are contracted in bf16 with fp32 accumulation whatever the input dtype -- fp32 / fp16 inputs
    are ROUNDED TO bf16 on entry (the reference scores them in their own dtype).  For bf16 inputs (what the Col* models
    produce) the fp32 scores here are more accurate than the reference's bf16-rounded ones; for fp32 inputs expect
    ~1e-2 relative difference from the reference, inside north_star's bf16 tolerance.
    """
    if len(qs) == 0:
        raise ValueError("No queries provided")
    if isinstance(ps, DocBank):  # a resident bank fixes the device
        if device is not None and _resolve_device(device) != ps.device:
            raise ValueError(f"the DocBank lives on {ps.device}, not on {device}")
        dev, bank = ps.device, ps
    else:
        if len(ps) == 0:
            raise ValueError("No passages provided")
        dev = _resolve_device(device)
        _require_cuda(dev)
        bank = DocBank.from_passages(ps, dev, batch_size=batch_size)
    q = QueryBlock(qs, dev)
    scores = maxsim(q, bank, round_bf16=round_bf16)
    out = scores.cpu()  # the reference contract: scores live on the CPU (:180)
    assert out.shape[0] == len(qs), f"Expected {len(qs)} scores, got {out.shape[0]}"
    return out.to(torch.float32)
