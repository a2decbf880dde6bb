"""GPU: corpus-sharded scoring runs the fused kernel per shard; 2-rank NCCL run when two GPUs are visible."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import colpali_b200 as cb
from colpali_b200.sharded import score_sharded, shard_bounds
from oracle import li_oracle as O

pytestmark = pytest.mark.gpu


def _corpus(n_docs):
    g = torch.Generator().manual_seed(0)
    qs = O.unit_rows((6, 32, 128), 1)
    docs = [O.unit_rows((int(n), 128), 100 + j) for j, n in enumerate(torch.randint(40, 400, (n_docs,), generator=g))]
    return qs, docs


def test_single_rank_matches_plain_scorer_and_topk():
    dev = torch.device("cuda:0")
    qs, docs = _corpus(37)
    bank = cb.DocBank.from_passages(docs, dev, reference_padding=False)
    full = score_sharded(qs.to(dev), bank, 0, 37)
    want = torch.from_numpy(O.maxsim_f64(list(qs), docs)).float()
    assert torch.allclose(full.cpu(), want, rtol=1e-5, atol=1e-4)
    ts, ti = score_sharded(qs.to(dev), bank, 0, 37, top_k=10)
    assert torch.equal(ti.cpu(), torch.topk(full.cpu(), 10, dim=1).indices)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    qs, docs = _corpus(41)
    lo, hi = shard_bounds(41, world)[rank]
    bank = cb.DocBank.from_passages(docs[lo:hi], dev, reference_padding=False)
    full = score_sharded(qs.to(dev), bank, lo, 41)
    ts, ti = score_sharded(qs.to(dev), bank, lo, 41, top_k=10)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), full=full.cpu().numpy(), ti=ti.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_nccl(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    qs, docs = _corpus(41)
    want = torch.from_numpy(O.maxsim_f64(list(qs), docs)).float()
    for r in range(2):
        got = np.load(os.path.join(tmp_path, f"r{r}.npz"))
        assert np.allclose(got["full"], want.numpy(), rtol=1e-5, atol=1e-4)
        assert np.array_equal(got["ti"], torch.topk(torch.from_numpy(got["full"]), 10, dim=1).indices.numpy())


def _fused_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from colpali_b200.sharded import FusedGatherScorer

    n_launch = 50
    docs = O.unit_rows((60, 200, 128), 50 + rank)          # every rank owns a different dense shard
    bank = cb.DocBank.from_passages(docs.to(dev), dev)
    qbs = [cb.QueryBlock(O.unit_rows((8, 32, 128), 500 + i).to(dev), dev) for i in range(n_launch)]  # same on all ranks
    ok = FusedGatherScorer.available(dev)
    res = {"available": np.array(ok)}
    if ok:
        # what every launch must produce: an NCCL all-gather of the local score slabs
        want = torch.empty(n_launch, world, 8, 60, device=dev)
        for i, qb in enumerate(qbs):
            dist.all_gather_into_tensor(want[i].view(world * 8, 60), cb.maxsim(qb, bank))
        for mc in (True, False):
            sc = FusedGatherScorer(8, 60, dev, use_multicast=mc)
            got = torch.empty_like(want)
            # 50 back-to-back launches with different queries, NO host synchronisation or barrier in between: the double
            # buffer + the per-launch completion counters are the only protection against overwriting unread slabs
            for i, qb in enumerate(qbs):
                view = sc.score(qb, bank)
                sc.wait()
                got[i].copy_(view)  # the consumer of launch i, stream-ordered before launch i + 2
            torch.cuda.synchronize()
            sc.check_status()
            res[f"equal_mc{int(mc)}"] = np.array([bool(torch.equal(got[i], want[i])) for i in range(n_launch)])
            res[f"used_mc{int(mc)}"] = np.array(sc.mc_base != 0)
            dist.barrier()
    np.savez(os.path.join(out_dir, f"f{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_allgather_two_ranks_back_to_back(tmp_path):
    """The kernel stores its scores straight into both ranks' gathered buffers (NVSwitch multicast or NVLink peer
    stores); 50 launches back to back, every one compared with an NCCL all-gather."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_fused_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(tmp_path, f"f{k}.npz")) for k in range(2)]
    if not bool(r[0]["available"]):
        pytest.skip("symmetric memory is not available in this environment")
    for k in range(2):
        for mc in (0, 1):
            assert r[k][f"equal_mc{mc}"].all(), (k, mc, np.nonzero(~r[k][f"equal_mc{mc}"])[0][:10])
        assert not bool(r[k]["used_mc0"])
