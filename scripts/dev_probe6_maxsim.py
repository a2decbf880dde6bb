"""Developer probe (GPU): how long may the epilogue hold an accumulator before the MMA stream stalls?"""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
q = F.normalize(torch.randn(32, 32, 128, device=dev), dim=-1).bfloat16()
d = F.normalize(torch.randn(1000, 1024, 128, device=dev), dim=-1).bfloat16()
bank = cb.DocBank.from_passages(d, dev); qb = cb.QueryBlock(q, dev)
for dbg, name in ((0x30000, "noTMA+noEpi"), (0x10000, "noEpi")):
  for delay in (0, 100, 200, 300, 400, 500, 600, 700, 800):
    _lib.set_option("debug_delay", delay)
    _lib.set_option("debug_flags", 0x40000 | dbg)
    for _ in range(4): s = cb.maxsim(qb, bank)
    torch.cuda.synchronize()
    f = s.flatten().double(); tot = f[:296].view(148, 2); x = f[512:512 + 8 * 148].view(148, 8); jobs = x[:, 7]
    print(f"{name:12s} delay {delay:4d}: cycles/job {(tot[:,0]/jobs).mean():6.0f} | issuer blocked epi {(x[:,1]/jobs).mean():4.0f} | epilogue wait {(x[:,2]/jobs).mean():4.0f} hold {(x[:,3]/jobs).mean():4.0f}", flush=True)
_lib.set_option("debug_delay", 0); _lib.set_option("debug_flags", 0)
