"""GPU, 2 ranks: the training exchange -- collective path (NCCL all-gather / reduce-scatter around the fused loss) and the
collective-free ``FusedExchange`` (push kernel, in-kernel wait, peer-scatter dD) -- against a single-process evaluation
of what the reference's trainers compute (contrastive_trainer.py:143-160, colmodel_torch_training.py:155-184):
loss_r = loss_fn(q_r, all_gather(pad(docs)), offset = r * B) on every rank, and through the gather's backward every rank
receives d(sum_r loss_r) / d(its own documents)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import colpali_b200 as cb
from colpali_b200 import exchange as X
from oracle import li_oracle as O

pytestmark = pytest.mark.gpu
B, LENS = 6, (180, 230)


def _inputs(rank):
    q = O.unit_rows((B, 20, 128), 700 + rank)
    d = O.unit_rows((B, LENS[rank], 128), 800 + rank)
    d[0, :3] = 0  # a document with masked (zero) rows
    return q, d


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    q, d = _inputs(rank)
    res = {}
    ex = None
    try:
        ex = X.FusedExchange(B, 256, dev)
        res["fused_available"] = np.array(True)
    except Exception as e:  # noqa: BLE001
        res["fused_available"] = np.array(False)
        res["why"] = np.array(repr(e)[:200])
    for lname, mod in (("colbert", cb.ColbertLoss()), ("pairwise", cb.ColbertPairwiseCELoss(pos_aware_negative_filtering=True))):
        for pad_first in (True, False):
            for path in ("nccl", "fused"):
                if path == "fused" and ex is None:
                    continue
                for rep in range(2 if path == "fused" else 1):  # twice: the symmetric buffers are reused across steps
                    qq, dd = q.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
                    loss = X.compute_loss_from_outputs(mod, qq, dd, pad_first=pad_first, fused=ex if path == "fused" else None)
                    loss.backward()
                    torch.cuda.synchronize()
                key = f"{lname}_{int(pad_first)}_{path}"
                res[key + "_loss"] = loss.detach().float().cpu().numpy()
                res[key + "_dq"] = qq.grad.float().cpu().numpy()
                res[key + "_dd"] = dd.grad.float().cpu().numpy()
    if ex is not None:
        ex.status_ok = int(ex.status.item()) == 0
        res["status_ok"] = np.array(ex.status_ok)
    np.savez(os.path.join(out_dir, f"x{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_exchange_two_ranks_against_single_process_oracle(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(os.path.join(tmp_path, f"x{k}.npz")) for k in range(2)]
    l_max = max(LENS)
    ports = {"colbert": lambda q, d, off: O.colbert_loss_port(q, d, offset=off),
             "pairwise": lambda q, d, off: O.colbert_pairwise_ce_loss_port(q, d, offset=off, pos_aware_negative_filtering=True)}
    for lname, port_fn in ports.items():
        for pad_first in (True, False):
            qs = [_inputs(k)[0].float().requires_grad_(True) for k in range(2)]
            ds = [_inputs(k)[1].float().requires_grad_(True) for k in range(2)]
            padded = []
            for d in ds:
                z = d.new_zeros(B, l_max - d.shape[1], 128)
                padded.append(torch.cat([z, d] if pad_first else [d, z], dim=1))
            gathered = torch.cat(padded, 0)
            losses = [port_fn(qs[k], gathered, k * B) for k in range(2)]
            sum(losses).backward()
            for path in ("nccl", "fused"):
                if path == "fused" and not bool(r[0]["fused_available"]):
                    continue
                for k in range(2):
                    key = f"{lname}_{int(pad_first)}_{path}"
                    assert abs(float(r[k][key + "_loss"]) - float(losses[k])) < 1e-3, (key, k)
                    for got, want in ((r[k][key + "_dq"], qs[k].grad), (r[k][key + "_dd"], ds[k].grad)):
                        real = want.abs().sum(-1) > 0  # zero (padding) rows: amax splits ties, the kernel does not
                        got = torch.from_numpy(got)
                        assert torch.allclose(got[real], want[real], rtol=2e-2, atol=want.abs().max().item() * 1e-2), (key, k)
    if bool(r[0]["fused_available"]):
        assert bool(r[0]["status_ok"]) and bool(r[1]["status_ok"])
