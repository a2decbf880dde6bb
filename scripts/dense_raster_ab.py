"""(python scripts/dense_raster_ab.py)  score_single_vector 1000 x 100000 x 1536 fp32 under different tile walks
(cpb_set_option "dense_raster"), interleaved; one JSON line."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(1000, 1536, generator=g, device=dev)
b = torch.randn(100000, 1536, generator=g, device=dev)
out = torch.empty(1000, 100000, device=dev)
groups = [1, 2, 4, 8, 16, 1 << 20]
ts = {k: [] for k in groups}
for k in groups:
    _lib.set_option("dense_raster", k)
    cb.dense_dot(a, b, out=out)
torch.cuda.synchronize()
for _ in range(6):
    for k in groups:
        _lib.set_option("dense_raster", k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            cb.dense_dot(a, b, out=out)
        e1.record(); torch.cuda.synchronize()
        ts[k].append(e0.elapsed_time(e1) / 3)
tr = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.mm(a, b.t(), out=out); e1.record(); torch.cuda.synchronize(); tr.append(e0.elapsed_time(e1))
print(json.dumps({"what": "dense 128-tile, 1000 x 100000 x 1536 fp32, median ms by raster group (1 = row tiles fastest, 2^20 = column tiles fastest)",
                  "ms": {str(k): sorted(v)[len(v) // 2] for k, v in ts.items()}, "torch_mm_ms": sorted(tr)[len(tr) // 2]}))
