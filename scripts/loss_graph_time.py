"""GPU time of the loss step at cfg3 without the Python launch path in the way: the forward (argmax mode + fused loss) and
forward + backward are captured into CUDA graphs of 20 steps and replayed.  One JSON line per variant."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from oracle import li_oracle as O

dev = torch.device("cuda:0")
q, d, _ = O.cfg3_inputs()
q, d = q.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
STEPS, REPLAYS = 20, 20


def time_graph(name, mod, backward):
    def step():
        if backward:
            q.grad = None; d.grad = None
            mod(q, d).backward()
        else:
            mod(q, d)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(STEPS): step()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPLAYS): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"what": name, "gpu_us_per_step": 1e3 * e0.elapsed_time(e1) / (STEPS * REPLAYS)}), flush=True)


for nm, mod in (("ColbertLoss", cb.ColbertLoss()), ("ColbertPairwiseCELoss", cb.ColbertPairwiseCELoss())):
    time_graph(f"{nm} forward (training: argmax + fused loss), CUDA graph", mod, False)
    time_graph(f"{nm} forward + backward, CUDA graph", mod, True)
time_graph("ColbertLoss(use_smooth_max) forward + backward, CUDA graph", cb.ColbertLoss(use_smooth_max=True), True)
