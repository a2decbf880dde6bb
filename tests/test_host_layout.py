"""CPU: host-side layout logic of the scorer (query padding, document bank metadata, the reference's zero-padding
floors) exercised on CPU tensors with the CUDA requirement patched out -- no kernel is called."""
import math

import pytest
import torch

from colpali_b200 import scoring
from oracle import li_oracle as O

CPU = torch.device("cpu")


@pytest.fixture(autouse=True)
def _allow_cpu(monkeypatch):
    monkeypatch.setattr(scoring, "_require_cuda", lambda device: None)


def test_query_block_pads_every_query_to_a_multiple_of_32_rows():
    qs = [torch.randn(5, 128), torch.randn(33, 128), torch.randn(1, 128)]
    qb = scoring.QueryBlock(qs, CPU)
    assert (qb.n, qb.nq_pad) == (3, 64) and qb.flat.shape == (3 * 64, 128) and qb.flat.dtype == torch.bfloat16
    blocks = qb.flat.view(3, 64, 128)
    for i, q in enumerate(qs):
        assert torch.equal(blocks[i, : q.shape[0]], q.bfloat16())
        assert not blocks[i, q.shape[0] :].any()          # zero rows contribute exactly 0 (processing_utils.py:172)
    # small embedding dims are zero-padded to 128 columns
    qb32 = scoring.QueryBlock(torch.randn(2, 4, 32), CPU)
    assert qb32.flat.shape == (2 * 32, 128) and not qb32.flat[:, 32:].any()
    # a tensor already in kernel layout is used without a copy
    q = torch.randn(4, 32, 128).bfloat16()
    assert scoring.QueryBlock(q, CPU).flat.data_ptr() == q.data_ptr()
    with pytest.raises(ValueError, match="No queries"):
        scoring.QueryBlock([], CPU)


def test_doc_bank_from_list_matches_reference_padding_groups():
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(0, 50, (300,), generator=g).tolist()
    ps = [torch.randn(n, 128) for n in lens]
    for bs in (128, 7):
        bank = scoring.DocBank.from_passages(ps, CPU, batch_size=bs)
        assert bank.n_docs == 300 and bank.contiguous and bank.max_len == max(lens) and bank.uniform_len == 0
        assert bank.length.tolist() == lens
        assert bank.start.tolist() == [sum(lens[:j]) for j in range(300)]
        assert bank.flat.shape == (sum(lens), 128)
        want = O.reference_floors(lens, bs)   # 0 where processing_utils.py:176-178 would zero-pad the document
        got = bank.floor.tolist()
        assert all((a == b) or (math.isinf(a) and math.isinf(b)) for a, b in zip(got, want))
    # equal lengths: the reference pads nothing -> no floor array at all
    same = scoring.DocBank.from_passages([torch.randn(9, 128) for _ in range(5)], CPU)
    assert same.floor is None and same.uniform_len == 9 and same.max_len == 9


def test_doc_bank_from_dense_tensor_is_zero_copy_and_uniform():
    d = torch.randn(6, 11, 128).bfloat16()
    bank = scoring.DocBank.from_passages(d, CPU)
    assert bank.flat.data_ptr() == d.data_ptr() and bank.flat.shape == (66, 128)
    assert bank.start.tolist() == [0, 11, 22, 33, 44, 55] and bank.length.tolist() == [11] * 6
    assert bank.floor is None and bank.contiguous and bank.uniform_len == 11 and bank.max_len == 11
    with pytest.raises(ValueError, match="No passages"):
        scoring.DocBank.from_passages([], CPU)
    with pytest.raises(scoring._lib.ColpaliB200Error):
        scoring.DocBank.from_passages(torch.randn(2, 3, 400), CPU)   # embedding dim > 320
    wide = scoring.DocBank.from_passages(torch.randn(2, 3, 300), CPU)  # padded to the next multiple of 64
    assert wide.flat.shape == (6, 320) and not wide.flat[:, 300:].any()
