"""Corpus-sharded late-interaction scoring over several B200s (one process per GPU, torch.distributed).

The reference scores on a single device (processing_utils.py:132-187 never touches torch.distributed); this is the
multi-GPU form BASELINE.json's configs[3] asks for.  Documents are independent units: the max over a document's
tokens and the sum over a query's tokens both complete on the GPU that owns the document, so the only exchange is
one all-gather of per-shard results over NVLink:

* full slabs  -- every rank contributes its ``[n_queries, n_local]`` fp32 scores (padded to the largest shard
                 with -inf), all ranks end with ``[n_queries, n_docs_total]``;
* top-k       -- every rank contributes its local top-k ``(score, global doc id)`` per query (KBs), all ranks
                 merge to the global top-k.  Ties are broken by the smaller document id, deterministically.  The local
                 selection is fused into the scoring kernel's tail (``scoring.maxsim_topk``, csrc/topk_tail.cuh) for
                 k <= 16, dim 128 and queries of at most 32 tokens; ``torch.topk`` on the slab otherwise.

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


def _local_topk(qs, bank, k: int, scorer: Optional[Callable]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(scores [n_q, k'], local document index [n_q, k'] int64, n_local), k' = min(k, n_local); order: score desc, index asc."""
    if scorer is None:
        from .scoring import QueryBlock, fused_topk_supported, maxsim_topk

        q = QueryBlock(qs, bank.device)
        if fused_topk_supported(q, bank, k):
            _, s, i = maxsim_topk(q, bank, k)
            return s, i.to(torch.int64), bank.n_docs
        local = _default_local_scorer(qs, bank)
    else:
        local = scorer(qs, bank).to(torch.float32)
    n_local = local.shape[1]
    ids = torch.arange(n_local, device=local.device).expand(local.shape[0], n_local)
    return (*merge_topk(local, ids, min(k, n_local)), n_local)


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
    if top_k is not None:
        top_k = min(top_k, n_docs_total)  # never return (-inf, INT64_MAX) filler candidates
        s, i, n_local = _local_topk(qs, local_bank, top_k, local_scorer)
        nq, dev = s.shape[0], s.device
    else:
        local = (local_scorer or _default_local_scorer)(qs, local_bank).to(torch.float32)
        nq, n_local = local.shape
    if world == 1:
        return local if top_k is None else (s, i + doc_offset)

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

    k_local = s.shape[1]
    cand_s = torch.full((nq, top_k), float("-inf"), dtype=torch.float32, device=dev)
    cand_i = torch.full((nq, top_k), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev)
    cand_s[:, :k_local] = s
    cand_i[:, :k_local] = i + doc_offset
    all_s = torch.empty(world, nq, top_k, dtype=torch.float32, device=dev)
    all_i = torch.empty(world, nq, top_k, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_s.view(world * nq, top_k), cand_s, group=group)
    dist.all_gather_into_tensor(all_i.view(world * nq, top_k), cand_i, group=group)
    return merge_topk(all_s.permute(1, 0, 2).reshape(nq, world * top_k), all_i.permute(1, 0, 2).reshape(nq, world * top_k), top_k)


class FusedGatherScorer:
    """Corpus-sharded scoring whose all-gather is fused into the MaxSim kernel (no collective kernel at all).

    Every rank owns ``n_local`` documents; the kernel epilogue stores each score into ALL ranks' copies of
    ``gathered[slab][my_rank]`` in a symmetric-memory buffer (``torch.distributed._symmetric_memory``) -- one
    ``multimem.st`` through the NVSwitch multicast mapping when the fabric offers one, else one store per NVLink peer
    mapping -- and every CTA, once its scores are out, adds 1 (release, system scope) to a per-(slab, rank) counter on
    every rank.  ``wait()`` enqueues a kernel that spins until the local counters have grown by every rank's grid size.

    The slab is triple-buffered: launch ``k + 3`` of a rank overwrites the slab of launch ``k``.  Before its first score
    store the kernel itself waits (on LOCAL counters) until every rank has completed launch ``k + 1`` -- hence started
    it, which by stream order on that rank is after it finished with the results of launch ``k`` (rule: enqueue the
    consumers of a result on the same stream before the next ``score``).  So back-to-back launches need no host round
    trip and no extra kernel, and launch ``k + 3`` never has to wait for launch ``k + 2``: with ``independent=True`` it
    may start while the previous launch is still draining its remote stores.

    Needs queries of at most 32 tokens and equally sized shards.  ``available()`` tells whether symmetric memory can be
    set up in this process group (agreed across ranks); callers fall back to ``score_sharded`` (NCCL all-gather) otherwise.
    """

    _SLABS = 3
    _FLAG_WORDS = 3 * 64  # [3 slabs][64 ranks]

    def __init__(self, n_queries: int, n_local: int, device: torch.device, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if self.world > 64:
            raise ValueError("FusedGatherScorer supports at most 64 ranks")
        self.n_queries, self.n_local, self.device = n_queries, n_local, device
        self.n_slab = self.world * n_queries * n_local          # words of one slab: gathered[world, n_q, n_local]
        self.buf = symm_mem.empty(self._SLABS * self.n_slab + self._FLAG_WORDS, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peer_ptrs = torch.tensor(list(self.hdl.buffer_ptrs), dtype=torch.int64, device=device)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        # all ranks must agree on the store path (a rank without the mapping would miss the others' multicast stores)
        agree = torch.tensor([1 if mc else 0], dtype=torch.int32, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        self.mc_base = mc if int(agree) else 0
        self.step = 0            # launches issued
        self.count = [0] * self._SLABS  # CTAs that have signalled per slab (every rank launches the same grid)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl.barrier()

    @staticmethod
    def available(device: torch.device, group=None) -> bool:
        """True on every rank or on none (the outcome of the local set-up is all-reduced)."""
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm_mem

            t = symm_mem.empty(64, dtype=torch.float32, device=device)
        except Exception:  # noqa: BLE001 - any failure means "use the NCCL path"
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if not int(flag):
            return False
        try:  # collective: entered by every rank, since every rank got here
            symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
        except Exception:  # noqa: BLE001
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        return bool(int(flag))

    def _slab(self, k: int) -> torch.Tensor:
        return self.buf[k * self.n_slab: (k + 1) * self.n_slab].view(self.world, self.n_queries, self.n_local)

    def _flags(self, k: int) -> torch.Tensor:
        return self.buf[self._SLABS * self.n_slab + k * 64:]

    def score(self, q, bank, independent: bool = False) -> torch.Tensor:
        """Enqueue the fused kernel; returns the local gathered view ``[world, n_queries, n_local]`` of this launch's
        slab (complete after ``wait()``).  ``independent``: see ``scoring.maxsim``."""
        from .scoring import launch_maxsim

        if q.n != self.n_queries or bank.n_docs != self.n_local or q.nq_pad != 32:
            raise ValueError("FusedGatherScorer was built for a different shape (or queries longer than 32 tokens)")
        k = self.step % self._SLABS
        gather = {
            "peer_bases": self.peer_ptrs.data_ptr(), "mc_base": self.mc_base, "n_peers": self.world,
            "slab_word_offset": k * self.n_slab + self.rank * self.n_queries * self.n_local,
            "flag_word_offset": self._SLABS * self.n_slab + k * 64 + self.rank,
        }
        if self.step >= 2:  # write-after-read guard: launch step - 2 (slab k + 1) is complete on every rank
            g = (k + 1) % self._SLABS
            gather["wait_flags"] = self._flags(g).data_ptr()
            gather["wait_value"] = self.count[g] & 0xFFFFFFFF
        grid = launch_maxsim(q, bank, scores=None, gather=gather, independent=independent)
        self.count[k] += grid  # every rank launches the same shape, hence the same grid
        self.step += 1
        return self._slab(k)

    def wait(self) -> None:
        """Enqueue a wait until every rank has signalled completion of its most recent ``score`` launch (all ranks must
        have issued the same number of launches): afterwards the gathered view of that launch is complete."""
        from . import _lib

        if self.step == 0:
            return
        k = (self.step - 1) % self._SLABS
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.cpb_wait_flags(self._flags(k).data_ptr(), self.world, self.count[k] & 0xFFFFFFFF,
                                    self.status.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "cpb_wait_flags")
        _lib.count_launches(1)

    def check_status(self) -> None:
        """Synchronising: raise if a wait timed out (a peer never signalled)."""
        late = int(self.status.item())
        if late:
            raise RuntimeError(f"fused all-gather: ranks with bit mask {late:#x} did not signal within the time-out")

    def barrier(self) -> None:
        """Device-side cross-rank barrier on the current stream (symmetric-memory signal pads)."""
        self.hdl.barrier()
