"""Timings of the wide-embedding kernels: K-pipelined scorer at dims 192/256/320 (cfg2 shape), wide head at dims 256/320 with
and without the 2-CTA W multicast.  One JSON line each."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
from oracle import li_oracle as O

dev = torch.device("cuda:0")
try:
    PEAKS = json.load(open("MEASURED_PEAKS.json"))
except Exception:
    PEAKS = {}
PEAK_HBM = PEAKS.get("hbm_gbs", 6569.6)
PEAK_TF = PEAKS.get("bf16_tflops", 1500.0)


def cuda_time(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for dim in (128, 192, 256, 320):
    q = O.unit_rows((32, 32, dim), 1).to(dev)
    d = O.unit_rows((1000, 1030, dim), 2).to(dev)
    qb, bank = cb.QueryBlock(q, dev), cb.DocBank.from_passages(d, dev)
    t = cuda_time(lambda: cb.maxsim(qb, bank))
    fl = 2 * 32 * 32 * 1000 * 1030 * dim
    print(json.dumps({"what": f"maxsim cfg2 shape, dim={dim}", "ms": t, "tflops": fl / t / 1e9, "frac_tensor": fl / t / 1e9 / PEAK_TF,
                      "queries_per_s": 32 / t * 1e3}), flush=True)

for tokens, hidden, dim in ((275 * 1000, 2560, 320), (275 * 1000, 2048, 320), (275 * 1000, 1536, 256)):
    h = torch.randn(tokens, hidden, device=dev).bfloat16()
    lin = torch.nn.Linear(hidden, dim).to(dev, torch.bfloat16)
    mask = torch.ones(tokens, dtype=torch.long, device=dev)
    for cl in (1, 2):
        _lib.set_option("head_cluster", cl)
        t = cuda_time(lambda: cb.fused_head(h, lin.weight, lin.bias, mask))
        by = 2 * tokens * (hidden + dim) + 2 * hidden * dim
        fl = 2 * tokens * hidden * dim
        print(json.dumps({"what": f"fused_head T={tokens} H={hidden} dim={dim} cluster={cl}", "ms": t, "gbs": by / t / 1e6,
                          "frac_hbm": by / t / 1e6 / PEAK_HBM, "tflops": fl / t / 1e9, "frac_tensor": fl / t / 1e9 / PEAK_TF}), flush=True)
    _lib.set_option("head_cluster", 0)
    t_ref = cuda_time(lambda: O.head_port(h, lin.weight, lin.bias, mask), n=5, warm=2)
    print(json.dumps({"what": f"reference head chain on the same GPU T={tokens} H={hidden} dim={dim}", "ms": t_ref}), flush=True)
