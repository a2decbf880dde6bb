"""Corpus-sharded late-interaction scoring over several B200s (one process per GPU, torch.distributed).

The reference scores on a single device (processing_utils.py:132-187 never touches torch.distributed); this is the
multi-GPU form BASELINE.json's configs[3] asks for.  Documents are independent units: the max over a document's
tokens and the sum over a query's tokens both complete on the GPU that owns the document, so the only exchange is
one all-gather of per-shard results over NVLink:

* full slabs  -- every rank contributes its ``[n_queries, n_local]`` fp32 scores (padded to the largest shard
                 with -inf), all ranks end with ``[n_queries, n_docs_total]``;
* top-k       -- every rank contributes its local top-k ``(score, global doc id)`` per query (KBs), all ranks
                 merge to the global top-k.  Ties are broken by the smaller document id, deterministically.

``local_scorer`` is the function that scores the local shard; it defaults to the fused sm_100a kernel and exists
so the host-side logic (sharding, padding, id offsets, merge) can be exercised under ``gloo`` on CPU in tests.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_docs: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced document ranges ``[lo, hi)`` per rank (the first ``n_docs % world`` ranks get one more)."""
    base, rem = divmod(n_docs, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _default_local_scorer(qs, bank) -> torch.Tensor:
    from .scoring import QueryBlock, maxsim

    return maxsim(QueryBlock(qs, bank.device), bank)


def _world(group) -> Tuple[int, int]:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def merge_topk(scores: torch.Tensor, ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global top-k of candidate lists ``[n_queries, m]``: larger score first, smaller document id on ties."""
    k = min(k, scores.shape[1])
    order = torch.argsort(ids, dim=1, stable=True)                      # secondary key: id ascending
    s, i = torch.gather(scores, 1, order), torch.gather(ids, 1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)       # primary key: score descending
    s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
    return s[:, :k].contiguous(), i[:, :k].contiguous()


def score_sharded(
    qs,
    local_bank,
    doc_offset: int,
    n_docs_total: int,
    *,
    top_k: Optional[int] = None,
    group=None,
    local_scorer: Optional[Callable] = None,
):
    """Score ``qs`` (replicated on every rank) against the corpus whose documents ``[doc_offset, doc_offset + n_local)``
    live in ``local_bank`` on this rank.

    Returns ``scores [n_queries, n_docs_total]`` (fp32, on the local device) when ``top_k`` is None, else
    ``(scores [n_queries, k], doc_ids [n_queries, k])`` -- identical on every rank.
    """
    rank, world = _world(group)
    scorer = local_scorer or _default_local_scorer
    local = scorer(qs, local_bank).to(torch.float32)
    nq, n_local = local.shape
    if world == 1:
        if top_k is None:
            return local
        ids = torch.arange(doc_offset, doc_offset + n_local, device=local.device).expand(nq, n_local)
        return merge_topk(local, ids, top_k)

    bounds = shard_bounds(n_docs_total, world)
    if (doc_offset, doc_offset + n_local) != bounds[rank]:
        raise ValueError(f"rank {rank} holds documents [{doc_offset}, {doc_offset + n_local}) but the balanced sharding "
                         f"of {n_docs_total} documents over {world} ranks assigns it {bounds[rank]}")
    if top_k is None:
        width = max(hi - lo for lo, hi in bounds)
        slab = torch.full((nq, width), float("-inf"), dtype=torch.float32, device=local.device)
        slab[:, :n_local] = local
        gathered = torch.empty(world, nq, width, dtype=torch.float32, device=local.device)
        dist.all_gather_into_tensor(gathered.view(world * nq, width), slab, group=group)
        return torch.cat([gathered[r, :, : hi - lo] for r, (lo, hi) in enumerate(bounds)], dim=1)

    k_local = min(top_k, n_local)
    s, i = torch.topk(local, k_local, dim=1)
    cand_s = torch.full((nq, top_k), float("-inf"), dtype=torch.float32, device=local.device)
    cand_i = torch.full((nq, top_k), torch.iinfo(torch.int64).max, dtype=torch.int64, device=local.device)
    cand_s[:, :k_local] = s
    cand_i[:, :k_local] = i + doc_offset
    all_s = torch.empty(world, nq, top_k, dtype=torch.float32, device=local.device)
    all_i = torch.empty(world, nq, top_k, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(all_s.view(world * nq, top_k), cand_s, group=group)
    dist.all_gather_into_tensor(all_i.view(world * nq, top_k), cand_i, group=group)
    return merge_topk(all_s.permute(1, 0, 2).reshape(nq, world * top_k), all_i.permute(1, 0, 2).reshape(nq, world * top_k), top_k)
