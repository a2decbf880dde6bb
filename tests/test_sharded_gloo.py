"""CPU: host-side logic of corpus-sharded scoring under gloo with world_size 2 and 3 (the local scorer is the
oracle here; on the GPU box the same code path runs the fused kernel -- tests/test_sharded_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colpali_b200.sharded import merge_topk, score_sharded, shard_bounds
from oracle import li_oracle as O


def test_shard_bounds():
    assert shard_bounds(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for n, w in ((100000, 8), (7, 2), (0, 2)):
        b = shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_merge_topk_ties_by_smaller_id():
    s = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]])
    i = torch.tensor([[9, 7, 2, 5, 4]])
    ts, ti = merge_topk(s, i, 3)
    assert ts.tolist() == [[3.0, 3.0, 3.0]] and ti.tolist() == [[2, 4, 7]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_scorer(qs, bank):  # bank here is simply the list of this rank's documents
    return torch.from_numpy(O.maxsim_f64(list(qs), bank)).float()


def _worker(rank, world, port, n_docs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    qs = O.unit_rows((5, 8, 32), 1, torch.float32)
    docs = [O.unit_rows((int(n), 32), 100 + j, torch.float32) for j, n in enumerate(torch.randint(3, 20, (n_docs,), generator=g))]
    lo, hi = shard_bounds(n_docs, world)[rank]
    full = score_sharded(qs, docs[lo:hi], lo, n_docs, local_scorer=_oracle_scorer)
    ts, ti = score_sharded(qs, docs[lo:hi], lo, n_docs, top_k=4, local_scorer=_oracle_scorer)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), full=full.numpy(), ts=ts.numpy(), ti=ti.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_docs", [(2, 11), (3, 7)])
def test_sharded_equals_single_process(tmp_path, world, n_docs):
    mp.spawn(_worker, args=(world, _free_port(), n_docs, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    qs = O.unit_rows((5, 8, 32), 1, torch.float32)
    docs = [O.unit_rows((int(n), 32), 100 + j, torch.float32) for j, n in enumerate(torch.randint(3, 20, (n_docs,), generator=g))]
    want = torch.from_numpy(O.maxsim_f64(list(qs), docs)).float()
    want_s, want_i = torch.topk(want, 4, dim=1)
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"r{r}.npz"))
        assert np.array_equal(got["full"], want.numpy())           # every rank holds the full score matrix
        assert np.array_equal(got["ts"], want_s.numpy())
        assert np.array_equal(got["ti"], want_i.numpy())           # recall@k = 1 against the single-process pass


def test_merge_topk_and_shard_bounds_properties():
    """Randomised: merge_topk == a plain sort by (score desc, id asc), ties included (scores drawn from 4 values);
    shard_bounds is a partition into contiguous ranges whose sizes differ by at most one, larger shards first."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 4), st.integers(1, 40), st.integers(1, 45), st.integers(0, 2**31 - 1))
    def merge(nq, m, k, seed):
        g = torch.Generator().manual_seed(seed)
        s = torch.randint(0, 4, (nq, m), generator=g).float()
        s[torch.rand(nq, m, generator=g) < 0.1] = float("-inf")            # the filler of short shards
        ids = torch.stack([torch.randperm(1000, generator=g)[:m] for _ in range(nq)])
        ts, ti = merge_topk(s, ids, k)
        kk = min(k, m)
        assert ts.shape == ti.shape == (nq, kk)
        for r in range(nq):
            want = sorted(zip(s[r].tolist(), ids[r].tolist()), key=lambda p: (-p[0], p[1]))[:kk]
            assert list(zip(ts[r].tolist(), ti[r].tolist())) == want

    @settings(max_examples=100, deadline=None)
    @given(st.integers(0, 10**6), st.integers(1, 64))
    def bounds(n, w):
        b = shard_bounds(n, w)
        sizes = [hi - lo for lo, hi in b]
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)

    merge()
    bounds()
