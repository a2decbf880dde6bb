"""In-batch-negative exchange for multi-GPU training (one process per GPU, torch.distributed).

What the reference's trainers do around the loss when ``world_size > 1``:

* ``ContrastiveTrainer._compute_loss_from_outputs`` (colpali_engine/trainer/contrastive_trainer.py:137-160):
  ``accelerator.pad_across_processes(docs, dim=1, pad_index=0, pad_first=True)``, an autograd-aware all-gather
  (``concat_all_gather`` :14-17), then ``loss_func(q, gathered_docs, [neg], offset=rank * batch_size)``;
  explicit negatives stay local (:190-191).
* ``ColModelTorchTraining`` (trainer/colmodel_torch_training.py:115-117,160-183): right-pads to the global
  maximum length, ``all_gather_tensor_autograd``, same ``offset``.

This module is that exchange, backend-agnostic (``nccl`` on the GPU box, ``gloo`` in the CPU tests): queries stay
local, every rank's ``[B, L_r, D]`` document block is zero-padded to the global maximum length (zero rows are
ordinary padding tokens for the loss kernels: they score exactly 0, as in the reference) and gathered to
``[world * B, L_max, D]``; the backward of the gather is a reduce-scatter of ``dD`` (sum over ranks), so every rank
receives the gradient of ALL ranks' losses with respect to its own documents -- what the reference's functional
collectives produce.

The loss itself is whatever module is passed in (``colpali_b200.ColbertLoss`` & co. on the GPU; the tests inject the
CPU oracle to exercise the host logic).

``FusedExchange`` is the same exchange WITHOUT collective kernels, for the in-batch losses on NVLink-connected GPUs: a
push kernel writes every rank's padded block into all ranks' copies of the gathered bank (symmetric memory, NVSwitch
multicast stores when available), the loss kernel waits for the pushes in its prologue, and the backward adds each
document's gradient rows directly into its owner rank's accumulator over NVLink (csrc/exchange_sm100.cu,
csrc/loss_sm100.cu) -- no all-gather, no pad/cat pass, no reduce-scatter.

The torch-loop trainer also gathers the explicit negatives (colmodel_torch_training.py:175) and then calls the loss with
B local queries against world * B negative groups -- which the reference losses' ``einsum("bnd,blsd->blns")`` rejects
for world > 1 (late_interaction_losses.py:238), so that combination only runs at world size 1 there; like the HF
trainer (contrastive_trainer.py:190-191) this module keeps negatives local.
"""

from __future__ import annotations

import ctypes
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def _world(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


class _GatherRows(torch.autograd.Function):
    """``[B, ...] -> [world * B, ...]`` (rank-major); backward = reduce-scatter(sum) of the incoming gradient."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group):
        rank, world = _world(group)
        ctx.group, ctx.rank, ctx.world, ctx.rows = group, rank, world, x.shape[0]
        x = x.contiguous()
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x, group=group)
        return torch.cat(parts, dim=0)

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        grad = grad.contiguous()
        mine = torch.empty_like(grad[: ctx.rows])
        if dist.get_backend(ctx.group) == "nccl":
            dist.reduce_scatter_tensor(mine, grad, op=dist.ReduceOp.SUM, group=ctx.group)
        else:  # gloo has no reduce-scatter: all-reduce, keep my block
            total = grad.clone()
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=ctx.group)
            mine = total[ctx.rank * ctx.rows: (ctx.rank + 1) * ctx.rows].clone()
        return mine, None


def gather_with_grad(x: torch.Tensor, group=None) -> torch.Tensor:
    """Autograd-aware all-gather along dim 0 (``gather_with_grad`` of colmodel_torch_training.py:115-117 and
    ``concat_all_gather`` of contrastive_trainer.py:14-17).  Every rank must pass the same shape.  Identity when
    torch.distributed is not initialised or the world has one rank."""
    if _world(group)[1] == 1:
        return x
    return _GatherRows.apply(x, group)


def pad_across_processes(x: torch.Tensor, dim: int = 1, pad_first: bool = True, group=None) -> torch.Tensor:
    """Zero-pad ``dim`` to its maximum over all ranks (``accelerator.pad_across_processes(..., pad_index=0)`` as
    called at contrastive_trainer.py:143-145).  ``pad_first=True`` puts the zero rows in front (the HF-trainer path),
    ``False`` behind (colmodel_torch_training.py:160-170)."""
    rank, world = _world(group)
    if world == 1:
        return x
    n = torch.tensor([x.shape[dim]], dtype=torch.int64, device=x.device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX, group=group)
    extra = int(n.item()) - x.shape[dim]
    if extra == 0:
        return x
    shape = list(x.shape)
    shape[dim] = extra
    zeros = x.new_zeros(shape)
    return torch.cat([zeros, x] if pad_first else [x, zeros], dim=dim)


def gather_documents(doc_embeddings: torch.Tensor, pad_first: bool = True, group=None) -> Tuple[torch.Tensor, int]:
    """``[B, L_r, D]`` on every rank -> (``[world * B, L_max, D]``, offset of this rank's positives)."""
    rank, _ = _world(group)
    batch = doc_embeddings.shape[0]
    gathered = gather_with_grad(pad_across_processes(doc_embeddings, dim=1, pad_first=pad_first, group=group), group)
    return gathered, rank * batch


def compute_loss_from_outputs(loss_func: Callable, query_outputs: torch.Tensor, pos_target_outputs: torch.Tensor,
                              neg_target_outputs: Optional[torch.Tensor] = None, *, pad_first: bool = True,
                              group=None, gather: bool = True, fused: Optional["FusedExchange"] = None) -> torch.Tensor:
    """``ContrastiveTrainer._compute_loss_from_outputs`` (contrastive_trainer.py:137-160): gather the positives of all
    ranks, call the loss with ``offset = rank * batch_size``.  Explicit negatives are paired with the local queries
    and are not gathered (contrastive_trainer.py:190-191).

    ``gather=False`` mirrors the reference's ``accelerator.sync_gradients`` condition (:143): on gradient-accumulation
    micro-steps it scores against the local documents only, offset 0, no collective.  ``fused`` routes the exchange of
    an in-batch loss through ``FusedExchange`` (no collective kernels)."""
    if not gather or _world(group)[1] == 1:
        if neg_target_outputs is None:
            return loss_func(query_outputs, pos_target_outputs, offset=0)
        return loss_func(query_outputs, pos_target_outputs, neg_target_outputs, offset=0)
    if fused is not None and neg_target_outputs is None and fused.supports(loss_func, query_outputs, pos_target_outputs):
        return fused.loss(loss_func, query_outputs, pos_target_outputs, pad_first=pad_first)
    docs, offset = gather_documents(pos_target_outputs, pad_first=pad_first, group=group)
    if neg_target_outputs is None:
        return loss_func(query_outputs, docs, offset=offset)
    return loss_func(query_outputs, docs, neg_target_outputs, offset=offset)


class _FusedExchangeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, docs, ex, mode, temperature, normalize, filt, thr, factor, pad_first):
        from . import losses as L
        from .scoring import DocBank, QueryBlock

        dev = q.device
        b, l_r, dim = docs.shape
        l_max = ex.max_len_across_ranks(l_r)
        bank_view, wait = ex.push(docs.detach(), l_max, pad_first)      # [world * B, l_max, dim] on this rank
        qb = QueryBlock(q.detach(), dev)
        bank = DocBank.from_passages(bank_view, dev)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        c = bank.n_docs
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g = torch.empty(qb.n, c, dtype=torch.float32, device=dev) if need_grad else None
        desc = L._loss_desc(mode, temperature, normalize, filt, thr, factor, ex.rank * b, loss, g)
        _, aux = L._maxsim_for_loss(qb, bank, need_grad, 0.0, int(q.shape[1]), loss_desc=desc, wait=wait)
        if need_grad:
            ctx.save_for_backward(qb.flat, bank.flat, bank.start, bank.length, aux, g)
            ctx.meta = (qb.n, qb.nq_pad, int(q.shape[1]), c, l_max, l_r, pad_first, tuple(q.shape), q.dtype, docs.dtype)
            ctx.ex = ex
        return loss[0]

    @staticmethod
    def backward(ctx, grad_out):
        from . import losses as L

        q_flat, d_flat, d_start, d_len, aux, g = ctx.saved_tensors
        b, nq_pad, nq_real, c, l_max, l_r, pad_first, q_shape, q_dtype, d_dtype = ctx.meta
        ex = ctx.ex
        want_q, want_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        go = grad_out.detach().to(torch.float32).reshape(1).contiguous()
        dq, _ = L._maxsim_backward(g, go, aux, 0.0, nq_real, q_flat, b, nq_pad, d_flat, d_start, d_len, c, l_max,
                                   want_q, False)
        grad_q = dq.view(b, nq_pad, -1)[:, : q_shape[1], : q_shape[2]].to(q_dtype) if want_q else None
        # dD of ALL world * B documents: every document's rows are added into its owner's accumulator over NVLink.
        # Every rank must take part (the owners wait for world signals), so this runs even if docs need no gradient here.
        L._maxsim_backward(g, go, aux, 0.0, nq_real, q_flat, b, nq_pad, d_flat, d_start, d_len, c, l_max, False, True,
                           dd_doc_base=ex.dd_doc_base(l_max))
        acc = ex.finish_backward(l_max)                                   # [B, l_max, dim] fp32, complete
        grad_d = None
        if want_d:
            grad_d = (acc[:, l_max - l_r:] if pad_first else acc[:, :l_r]).to(d_dtype)
        return (grad_q, grad_d) + (None,) * 8


class FusedExchange:
    """Symmetric-memory state of the collective-free exchange for one (batch size, max length, dim) training setup.

    ``loss(loss_module, q, docs)`` == ``loss_module(q, all_gather(pad(docs)), offset=rank * B)`` including gradients,
    for ``ColbertLoss`` / ``ColbertPairwiseCELoss`` with the hard max, embedding dim 128, queries of at most 32 tokens.
    Per step and rank: one push kernel (forward), the fused MaxSim + loss kernel, dQ and dD kernels, one signal and one
    wait kernel (backward).  A device-side barrier (symmetric-memory signal pads) opens every step, so a rank never
    overwrites a bank or accumulator a slower peer is still reading.
    """

    def __init__(self, batch: int, max_len_cap: int, device: torch.device, group=None, dim: int = 128,
                 use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 64:
            raise ValueError("FusedExchange supports at most 64 ranks")
        self.batch, self.cap, self.dim, self.device = batch, max_len_cap, dim, device
        bank_words = self.world * batch * max_len_cap * dim // 2       # bf16 gathered bank, in 4-byte words
        acc_words = batch * max_len_cap * dim                          # fp32 accumulator of this rank's documents
        self.bank_off, self.acc_off, self.flag_off = 0, bank_words, bank_words + acc_words
        self.buf = symm_mem.empty(bank_words + acc_words + 128, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peer_ptrs_host = [int(p) for p in self.hdl.buffer_ptrs]
        self.peer_ptrs = torch.tensor(self.peer_ptrs_host, dtype=torch.int64, device=device)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        agree = torch.tensor([1 if mc else 0], dtype=torch.int32, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        self.mc_base = mc if int(agree) else 0
        self.push_count = 0      # CTAs every rank's push counter has received (all ranks push the same shape)
        self.bwd_count = 0       # backward signals (one per rank and step)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self._doc_base_cache: dict = {}
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl.barrier()

    # -- capability ------------------------------------------------------------------------------------------------
    def supports(self, loss_func, q: torch.Tensor, docs: torch.Tensor) -> bool:
        from .losses import ColbertLoss, ColbertPairwiseCELoss

        return (isinstance(loss_func, (ColbertLoss, ColbertPairwiseCELoss)) and not loss_func.use_smooth_max
                and q.dim() == 3 and docs.dim() == 3 and q.shape[1] <= 32 and q.shape[2] == self.dim == docs.shape[2]
                and docs.shape[0] == self.batch and docs.shape[1] <= self.cap and docs.dtype == torch.bfloat16)

    def max_len_across_ranks(self, l_r: int) -> int:
        n = torch.tensor([l_r], dtype=torch.int64, device=self.device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self.group)   # the reference syncs here too (.item(), :160-166)
        return int(n.item())

    # -- forward -----------------------------------------------------------------------------------------------------
    def push(self, docs: torch.Tensor, l_max: int, pad_first: bool):
        """Write this rank's padded block into every rank's gathered bank.  Returns (the LOCAL gathered view
        ``[world * B, l_max, dim]`` -- complete once the counters in ``wait`` are reached -- , wait tuple)."""
        from . import _lib

        lib = _lib.load()
        b, l_r, dim = docs.shape
        if not docs.is_contiguous():
            docs = docs.contiguous()
        self.hdl.barrier()  # every rank is done with the previous step's bank and accumulator
        acc = self.buf[self.acc_off: self.acc_off + b * l_max * dim]
        acc.zero_()         # ordered before this rank's push, whose counters every adder waits on
        a = _lib.ExchangePushArgs()
        a.pad_first = 1 if pad_first else 0
        a.d_src, a.n_docs, a.len, a.slot_len, a.dim = docs.data_ptr(), b, l_r, l_max, dim
        a.d_peer_bases, a.mc_base, a.n_peers = self.peer_ptrs.data_ptr(), self.mc_base, self.world
        a.bank_word_offset = self.bank_off + self.rank * b * l_max * dim // 2
        a.flag_word_offset = self.flag_off + self.rank
        with torch.cuda.device(self.device):
            a.stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = lib.cpb_exchange_push(ctypes.byref(a))
        _lib.check(rc, "cpb_exchange_push")
        _lib.count_launches(1)
        self.push_count += int(a.grid_out)
        words = self.world * b * l_max * dim // 2
        view = self.buf[self.bank_off: self.bank_off + words].view(torch.bfloat16).view(self.world * b, l_max, dim)
        flags = self.buf[self.flag_off:]
        return view, (flags.data_ptr(), self.world, self.push_count & 0xFFFFFFFF)

    def loss(self, loss_func, q: torch.Tensor, docs: torch.Tensor, pad_first: bool = True) -> torch.Tensor:
        from . import _lib

        mode = _lib.CPB_LOSS_CE if type(loss_func).__name__ == "ColbertLoss" else _lib.CPB_LOSS_PAIRWISE
        return _FusedExchangeLossFn.apply(q, docs, self, mode, loss_func.temperature, loss_func.normalize_scores,
                                          loss_func.pos_aware_negative_filtering, loss_func.filter_threshold,
                                          loss_func.filter_factor, pad_first)

    # -- backward ----------------------------------------------------------------------------------------------------
    def dd_doc_base(self, l_max: int) -> torch.Tensor:
        """int64 [world * B]: address of document c's [l_max, dim] fp32 block in its owner's accumulator."""
        t = self._doc_base_cache.get(l_max)
        if t is None:
            stride = l_max * self.dim * 4
            addrs = [self.peer_ptrs_host[c // self.batch] + 4 * self.acc_off + (c % self.batch) * stride
                     for c in range(self.world * self.batch)]
            t = torch.tensor(addrs, dtype=torch.int64, device=self.device)
            if len(self._doc_base_cache) > 16:
                self._doc_base_cache.clear()
            self._doc_base_cache[l_max] = t
        return t

    def finish_backward(self, l_max: int) -> torch.Tensor:
        """Publish this rank's gradient adds, wait for everyone's, return the local accumulator ``[B, l_max, dim]``."""
        from . import _lib

        lib = _lib.load()
        self.bwd_count += 1
        flags = self.buf[self.flag_off + 64:]
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = lib.cpb_signal_peers(self.peer_ptrs.data_ptr(), self.mc_base, self.world, self.flag_off + 64 + self.rank, stream)
            _lib.check(rc, "cpb_signal_peers")
            rc = lib.cpb_wait_flags(flags.data_ptr(), self.world, self.bwd_count & 0xFFFFFFFF, self.status.data_ptr(), stream)
            _lib.check(rc, "cpb_wait_flags")
        _lib.count_launches(2)
        return self.buf[self.acc_off: self.acc_off + self.batch * l_max * self.dim].view(self.batch, l_max, self.dim)
