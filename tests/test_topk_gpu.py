"""GPU: the top-k selection fused into the MaxSim kernel's tail (csrc/topk_tail.cuh) against a sort of the score
matrix the same launch wrote: larger score first, smaller document index on ties."""
import pytest
import torch

import colpali_b200 as cb
from colpali_b200.scoring import DocBank, QueryBlock, fused_topk_supported, maxsim, maxsim_topk
from colpali_b200.sharded import merge_topk, score_sharded

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _unit(gen, *shape):
    return torch.nn.functional.normalize(torch.randn(*shape, generator=gen, device=DEV), dim=-1).bfloat16()


def _expect(scores, k):
    ids = torch.arange(scores.shape[1], device=scores.device).expand_as(scores)
    return merge_topk(scores, ids, k)


@pytest.mark.parametrize("n_q,n_docs,doc_len,k", [
    (32, 1000, 64, 10),     # one query group per cluster rank
    (128, 3000, 130, 10),   # cfg4 geometry: 16 query groups, tile-balanced partitions that cut documents
    (20, 257, 0, 16),       # ragged documents; 3 query groups on clusters of 2 (one padding group); k = CPB_TOPK_MAX
    (5, 7, 33, 10),         # fewer documents than k
    (1, 4000, 40, 1),
])
def test_fused_topk_equals_sorted_scores(n_q, n_docs, doc_len, k):
    gen = torch.Generator(device=DEV).manual_seed(n_q * 1000 + n_docs)
    qs = _unit(gen, n_q, 32, 128)
    if doc_len:
        ps = _unit(gen, n_docs, doc_len, 128)
        ps[3] = ps[1]                      # exact ties: equal scores must come out in document order
        if n_docs > 500:
            ps[499] = ps[1]
    else:
        lens = torch.randint(1, 300, (n_docs,), generator=torch.Generator().manual_seed(1)).tolist()
        ps = [_unit(gen, n, 128) for n in lens]
        ps[200] = ps[17].clone()
    bank = DocBank.from_passages(ps, DEV)
    q = QueryBlock(qs, DEV)
    assert fused_topk_supported(q, bank, k)
    for rep in range(3):                    # the per-group counters are reset by the kernel: repeated launches agree
        scores, top_s, top_i = maxsim_topk(q, bank, k)
        assert torch.equal(scores, maxsim(q, bank))
        want_s, want_i = _expect(scores, k)
        assert top_s.shape == want_s.shape == (n_q, min(k, n_docs))
        assert top_i.dtype == torch.int32 and torch.equal(top_s, want_s) and torch.equal(top_i.long(), want_i), rep
    if doc_len:
        hit = (top_i == 1).nonzero()
        for qi, pos in hit.tolist():        # the duplicate of document 1 directly follows it
            if pos + 1 < top_i.shape[1]:
                assert top_i[qi, pos + 1] == 3


def test_score_sharded_single_rank_uses_the_fused_selection():
    from colpali_b200 import _lib

    gen = torch.Generator(device=DEV).manual_seed(5)
    qs, ps = _unit(gen, 16, 32, 128), _unit(gen, 600, 70, 128)
    bank = DocBank.from_passages(ps, DEV)
    before = _lib.gpu_launches()
    s, i = score_sharded(qs, bank, doc_offset=1000, n_docs_total=600, top_k=10)
    assert _lib.gpu_launches() - before == 1          # one kernel: scores + selection
    want_s, want_i = _expect(maxsim(QueryBlock(qs, DEV), bank), 10)
    assert torch.equal(s, want_s) and torch.equal(i, want_i + 1000)
    # k above the in-kernel list length falls back to a sort of the slab (still on the GPU)
    s40, i40 = score_sharded(qs, bank, doc_offset=0, n_docs_total=600, top_k=40)
    want_s, want_i = _expect(maxsim(QueryBlock(qs, DEV), bank), 40)
    assert torch.equal(s40, want_s) and torch.equal(i40, want_i)
    assert not fused_topk_supported(QueryBlock(qs, DEV), bank, 40)
    with pytest.raises(cb.ColpaliB200Error):
        maxsim_topk(QueryBlock(qs, DEV), bank, 40)
