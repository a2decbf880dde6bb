"""Developer probe: how does the CPU reference path (oracle port, torch.einsum bf16) scale with threads on this host?"""
import os, sys, time
import torch
sys.path.insert(0, ".")
from oracle import li_oracle as O
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
print("torch threads default", torch.get_num_threads(), "interop", torch.get_num_interop_threads())
q, d = O.cfg2_inputs(128)
for t in (128, 64, 32, 16, 8, 4):
    if t > (os.cpu_count() or 1): continue
    torch.set_num_threads(t)
    O.score_multi_vector_port(q, d)
    t0 = time.perf_counter(); O.score_multi_vector_port(q, d); dt = time.perf_counter() - t0
    print(f"threads={t}: {dt*1e3:.1f} ms for 32q x 128 docs -> {32/(dt*1000/128):.1f} queries/s equivalent on 1000 docs", flush=True)
