"""Developer probe (GPU): balanced vs document-granular partitions at cfg2 (cycles of the slowest CTA and time)."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
q = F.normalize(torch.randn(32, 32, 128, device=dev), dim=-1).bfloat16()
d = F.normalize(torch.randn(1000, 1030, 128, device=dev), dim=-1).bfloat16()
bank = cb.DocBank.from_passages(d, dev); qb = cb.QueryBlock(q, dev)
for rep in range(3):
  for bal in (0, 1):
    _lib.set_option("balanced", bal)
    _lib.set_option("debug_flags", 0x40000)
    for _ in range(4): s = cb.maxsim(qb, bank)
    torch.cuda.synchronize()
    f = s.flatten().double(); tot = f[:296].view(148, 2); x = f[512:512 + 8 * 148].view(148, 8); jobs = x[:, 7]
    _lib.set_option("debug_flags", 0)
    for _ in range(5): cb.maxsim(qb, bank)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): cb.maxsim(qb, bank)
    e1.record(); torch.cuda.synchronize()
    print(f"balanced={bal}: CTA cycles max {tot[:,0].max():.0f} mean {tot[:,0].mean():.0f} min {tot[:,0].min():.0f} | jobs/CTA max {jobs.max():.0f} min {jobs.min():.0f} | {e0.elapsed_time(e1)*10:.1f} us/launch", flush=True)
_lib.set_option("balanced", 1)
