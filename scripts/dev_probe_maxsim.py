"""Developer probe (GPU): where does the time go?  debug flags: 0x10000 skip epilogue, 0x20000 no TMA, 0x40000 clocks."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
q = F.normalize(torch.randn(32, 32, 128, device=dev), dim=-1).bfloat16()
def timeit(qb, bank, n=50):
    for _ in range(5): cb.maxsim(qb, bank)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): s = cb.maxsim(qb, bank)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, s
cfgs = [(1, 2), (2, 2), (1, 1)] if len(sys.argv) < 2 else [(1, 2), (2, 2)]
for nd in (1024, 1030):
    d = F.normalize(torch.randn(1000, nd, 128, device=dev), dim=-1).bfloat16()
    bank = cb.DocBank.from_passages(d, dev); qb = cb.QueryBlock(q, dev)
    fl = 2 * 32 * 32 * 1000 * nd * 128
    for rep in range(2):
      for cluster, r in cfgs:
        _lib.set_option("cluster", cluster); _lib.set_option("qtiles_per_cta", r)
        for dbg, name in ((0, "normal"), (0x20000, "noTMA"), (0x10000, "noEpi")):
            _lib.set_option("debug_flags", dbg)
            ms, _ = timeit(qb, bank)
            _lib.set_option("debug_flags", dbg | 0x40000)
            for _ in range(3): s = cb.maxsim(qb, bank)
            torch.cuda.synchronize()
            v = s.flatten()[:296].view(148, 2).double()
            cyc, ns = v[:, 0], v[:, 1]
            print(f"Nd={nd} C={cluster} R={r} {name:7s}: {ms*1e3:6.1f} us {fl/ms/1e9:5.0f} TF/s | CTA cycles max {cyc.max():.0f} mean {cyc.mean():.0f} min {cyc.min():.0f} | clock {(cyc/ns).mean():.3f} GHz", flush=True)
_lib.set_option("cluster", 0); _lib.set_option("qtiles_per_cta", 0); _lib.set_option("debug_flags", 0)
