"""CPU: the multi-GPU training exchange (pad across ranks -> autograd all-gather -> loss with offset) under gloo with
world_size 2 and 3.  The loss is the CPU oracle here; on the GPU box the same host code wraps the fused loss modules."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colpali_b200 import exchange as X
from oracle import li_oracle as O

B, NQ, DIM = 3, 6, 16


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(rank):
    """Rank-dependent document length: the exchange has to pad across ranks."""
    q = O.unit_rows((B, NQ, DIM), 10 + rank, torch.float32)
    d = O.unit_rows((B, 9 + 4 * rank, DIM), 20 + rank, torch.float32)
    neg = O.unit_rows((B, 2, 7, DIM), 30 + rank, torch.float32)
    return q, d, neg


def _worker(rank, world, port, pad_first, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, d, neg = (t.requires_grad_(True) for t in _inputs(rank))
    loss = X.compute_loss_from_outputs(O.colbert_loss_port, q, d, pad_first=pad_first)
    loss.backward()
    out = dict(loss=loss.detach().numpy(), dq=q.grad.numpy().copy(), dd=d.grad.numpy().copy())
    q.grad = d.grad = None
    loss2 = X.compute_loss_from_outputs(O.colbert_negative_ce_loss_port, q, d, neg, pad_first=pad_first)
    loss2.backward()
    out.update(loss2=loss2.detach().numpy(), dq2=q.grad.numpy().copy(), dd2=d.grad.numpy().copy(), dn2=neg.grad.numpy().copy())
    docs, offset = X.gather_documents(d.detach(), pad_first=pad_first)
    out.update(shape=np.array(docs.shape), offset=np.array(offset))
    # gradient-accumulation micro-step (the reference gathers only when accelerator.sync_gradients, :143): local loss
    local = X.compute_loss_from_outputs(O.colbert_loss_port, q.detach(), d.detach(), pad_first=pad_first, gather=False)
    out.update(local=local.numpy(), local_ref=O.colbert_loss_port(q.detach(), d.detach(), offset=0).numpy())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,pad_first", [(2, True), (2, False), (3, True)])
def test_exchange_equals_single_process(tmp_path, world, pad_first):
    mp.spawn(_worker, args=(world, _free_port(), pad_first, str(tmp_path)), nprocs=world, join=True)
    ins = [_inputs(r) for r in range(world)]
    qs = [t[0].clone().requires_grad_(True) for t in ins]
    ds = [t[1].clone().requires_grad_(True) for t in ins]
    ns = [t[2].clone().requires_grad_(True) for t in ins]
    l_max = max(d.shape[1] for d in ds)

    def padded(d):
        z = d.new_zeros(d.shape[0], l_max - d.shape[1], d.shape[2])
        return torch.cat([z, d] if pad_first else [d, z], dim=1)

    allc = torch.cat([padded(d) for d in ds], dim=0)
    losses = [O.colbert_loss_port(qs[r], allc, offset=r * B) for r in range(world)]
    sum(losses).backward()  # every rank back-propagates its own loss; document gradients add up over ranks
    got = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    for r in range(world):
        assert tuple(got[r]["shape"]) == (world * B, l_max, DIM) and int(got[r]["offset"]) == r * B
        assert float(got[r]["local"]) == float(got[r]["local_ref"])  # gather=False: no exchange, offset 0
        assert abs(float(got[r]["loss"]) - float(losses[r].detach())) < 1e-6
        assert np.allclose(got[r]["dq"], qs[r].grad.numpy(), rtol=1e-5, atol=1e-7)
        assert np.allclose(got[r]["dd"], ds[r].grad.numpy(), rtol=1e-5, atol=1e-7)
    for t in qs + ds:
        t.grad = None
    allc = torch.cat([padded(d) for d in ds], dim=0)
    losses2 = [O.colbert_negative_ce_loss_port(qs[r], allc, ns[r], offset=r * B) for r in range(world)]
    sum(losses2).backward()
    for r in range(world):
        assert abs(float(got[r]["loss2"]) - float(losses2[r].detach())) < 1e-6
        assert np.allclose(got[r]["dq2"], qs[r].grad.numpy(), rtol=1e-5, atol=1e-7)
        assert np.allclose(got[r]["dd2"], ds[r].grad.numpy(), rtol=1e-5, atol=1e-7)
        assert np.allclose(got[r]["dn2"], ns[r].grad.numpy(), rtol=1e-5, atol=1e-7)


def test_single_process_is_identity():
    q, d, _ = _inputs(0)
    assert X.gather_with_grad(d) is d and X.pad_across_processes(d) is d
    docs, offset = X.gather_documents(d)
    assert docs is d and offset == 0
    assert float(X.compute_loss_from_outputs(O.colbert_loss_port, q, d)) == float(O.colbert_loss_port(q, d, offset=0))
