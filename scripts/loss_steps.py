"""A few forward + backward steps of the ColBERT losses at cfg3, for `ncu --metrics gpu__time_duration.sum` launch lists.
    python scripts/loss_steps.py [hard|smooth|neg] [steps]"""
import sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from oracle import li_oracle as O

kind = sys.argv[1] if len(sys.argv) > 1 else "hard"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
q, d, _ = O.cfg3_inputs()
q, d = q.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
mod = cb.ColbertLoss(use_smooth_max=(kind == "smooth"))
torch.cuda.synchronize()
for _ in range(steps):
    q.grad = None; d.grad = None
    torch.cuda.nvtx.range_push("step")
    mod(q, d).backward()
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("done", kind, steps)
