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
CPU oracle to exercise the host logic).  Fusing the gather into the scorer's TMA loads (peers' shards read straight
over NVLink) is the planned replacement for the collective; the interface here stays.
"""

from __future__ import annotations

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
                              group=None) -> torch.Tensor:
    """``ContrastiveTrainer._compute_loss_from_outputs`` (contrastive_trainer.py:137-160): gather the positives of all
    ranks, call the loss with ``offset = rank * batch_size``.  Explicit negatives are paired with the local queries
    and are not gathered (contrastive_trainer.py:190-191)."""
    docs, offset = gather_documents(pos_target_outputs, pad_first=pad_first, group=group)
    if neg_target_outputs is None:
        return loss_func(query_outputs, docs, offset=offset)
    return loss_func(query_outputs, docs, neg_target_outputs, offset=offset)
