"""Host-side cost of the loss module at cfg3: CUDA-event time of forward + backward per step, the same with the
launches only (no sync between steps), and a cProfile of the Python path.
    python scripts/host_profile.py [steps]"""
import cProfile, io, json, pstats, sys, time
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from oracle import li_oracle as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
q, d, _ = O.cfg3_inputs()
q, d = q.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
mod = cb.ColbertLoss()


def step():
    q.grad = None; d.grad = None
    mod(q, d).backward()


for _ in range(10): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps): step()
e1.record(); t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(json.dumps({"what": "ColbertLoss fwd+bwd at cfg3, back to back", "gpu_ms_per_step": e0.elapsed_time(e1) / steps,
                  "host_ms_per_step_enqueue": 1e3 * t_host / steps}), flush=True)
with torch.no_grad():
    for _ in range(10): mod(q, d)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for _ in range(steps): mod(q, d)
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
print(json.dumps({"what": "ColbertLoss forward only (no grad)", "gpu_ms_per_step": e0.elapsed_time(e1) / steps,
                  "host_ms_per_step_enqueue": 1e3 * t_host / steps}), flush=True)
pr = cProfile.Profile()
pr.enable()
for _ in range(steps): step()
pr.disable()
torch.cuda.synchronize()
buf = io.StringIO()
pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(35)
print(buf.getvalue()[:6000])
