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

from typing import Callable, List, Optional, Tuple

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

    top_k = min(top_k, n_docs_total)  # never return (-inf, INT64_MAX) filler candidates
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


class FusedGatherScorer:
    """Corpus-sharded scoring whose all-gather is fused into the MaxSim kernel (no collective kernel at all).

    Every rank owns ``n_local`` documents; the kernel epilogue stores each score into all ranks' copies of
    ``gathered[my_rank]`` through NVLink peer mappings of a symmetric-memory buffer
    (``torch.distributed._symmetric_memory``), and the last CTA of each rank's grid then stores a per-launch completion
    word into every peer; after ``wait()`` (or ``barrier()``) every rank holds ``[world, n_queries, n_local]``.
    Needs queries of at most 32 tokens and equally sized shards.  ``available()`` tells whether symmetric memory can be
    set up in this process group; callers fall back to ``score_sharded`` (NCCL all-gather) otherwise.
    """

    def __init__(self, n_queries: int, n_local: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.n_queries, self.n_local, self.device = n_queries, n_local, device
        self.n_slab = self.world * n_queries * n_local
        # gathered slabs followed by one completion word per rank (64 words reserved)
        self.buf = symm_mem.empty(self.n_slab + 64, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peer_ptrs = torch.tensor(list(self.hdl.buffer_ptrs), dtype=torch.int64, device=device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.step = 0
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl.barrier()

    @staticmethod
    def available(device: torch.device, group=None) -> bool:
        try:
            FusedGatherScorer(1, 1, device, group)
            return True
        except Exception:  # noqa: BLE001 - any failure means "use the NCCL path"
            return False

    def score(self, q, bank) -> torch.Tensor:
        """Enqueue the fused kernel; returns the local gathered view ``[world, n_queries, n_local]`` (the other ranks'
        slabs are complete after ``barrier()``)."""
        from . import _lib
        from .scoring import _EPOCH, _split_workspace

        if q.n != self.n_queries or bank.n_docs != self.n_local or q.nq_pad != 32:
            raise ValueError("FusedGatherScorer was built for a different shape (or queries longer than 32 tokens)")
        lib = _lib.load()
        dev = self.device
        flags = _lib.CPB_FLAG_CONTIGUOUS if bank.contiguous else 0
        split = None
        if bank.contiguous and bank.max_len > 0:
            split = _split_workspace(dev, lib.cpb_maxsim_split_workspace_bytes(q.n, q.nq_pad))
        _EPOCH[0] = _EPOCH[0] % 0xFFFFFFF0 + 1
        with torch.cuda.device(dev):
            rc = lib.cpb_maxsim_fwd_allgather(
                q.flat.data_ptr(), q.n, q.nq_pad, bank.flat.data_ptr(), bank.flat.shape[0],
                bank.start.data_ptr(), bank.length.data_ptr(),
                bank.floor.data_ptr() if bank.floor is not None else None, bank.n_docs,
                self.peer_ptrs.data_ptr(), self.world, self.rank, flags,
                bank.uniform_len, bank.max_len, split.data_ptr() if split is not None else None,
                split.numel() if split is not None else 0, _EPOCH[0],
                self.counter.data_ptr(), self.n_slab, self.step + 1, torch.cuda.current_stream(dev).cuda_stream,
            )
        _lib.check(rc, "cpb_maxsim_fwd_allgather")
        _lib.count_launches(1)
        self.step += 1
        return self.buf[: self.n_slab].view(self.world, self.n_queries, self.n_local)

    def wait(self) -> None:
        """Enqueue a wait until every rank has signalled completion of its most recent ``score`` launch (all ranks must
        have issued the same number of launches): afterwards the gathered view is complete."""
        from . import _lib

        lib = _lib.load()
        flags = self.buf[self.n_slab :]
        with torch.cuda.device(self.device):
            rc = lib.cpb_wait_flags(flags.data_ptr(), self.world, self.step,
                                    torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "cpb_wait_flags")
        _lib.count_launches(1)

    def barrier(self) -> None:
        """Device-side cross-rank barrier on the current stream (symmetric-memory signal pads): after it, every rank's
        kernel has finished and all slabs are visible."""
        self.hdl.barrier()
