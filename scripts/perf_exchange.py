"""Training exchange at cfg3 x N ranks (B = 64 pairs per rank, N_q = 32, documents 768..1030 tokens, C = 64 N): one
forward + backward of ColbertLoss through (a) the collective path of colpali_b200.exchange (NCCL all-gather of the padded
documents, reduce-scatter of dD) and (b) FusedExchange (push kernel, in-kernel wait, peer-scatter dD).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/perf_exchange.py
"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import colpali_b200 as cb
from colpali_b200 import exchange as X

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device=dev).manual_seed(10 + rank)
B, L = 64, 1030 - 37 * rank
q = torch.nn.functional.normalize(torch.randn(B, 32, 128, device=dev, generator=g), dim=-1).bfloat16().requires_grad_(True)
d = torch.nn.functional.normalize(torch.randn(B, L, 128, device=dev, generator=g), dim=-1).bfloat16().requires_grad_(True)
mod = cb.ColbertLoss()
ex = X.FusedExchange(B, 1030, dev)


def step(fused):
    q.grad = None; d.grad = None
    loss = X.compute_loss_from_outputs(mod, q, d, fused=ex if fused else None)
    loss.backward()
    return loss


out = {}
for name, fused in (("nccl", False), ("fused", True), ("nccl2", False), ("fused2", True)):
    for _ in range(3): loss = step(fused)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): loss = step(fused)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[name + "_ms"] = float(t); out[name + "_loss"] = float(loss)
    out[name + "_dd_checksum"] = float(d.grad.float().abs().sum())
if rank == 0:
    out.update(world=world, B=B, C=B * world, multicast=bool(ex.mc_base))
    print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
