"""Developer probe (GPU): is the MMA overhead per instruction or per job?  (debug flag 0x80000 doubles the MMAs of a job)"""
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
for name, dbg in (("normal", 0), ("noEpi+noTMA", 0x30000), ("noEpi+noTMA+doubleMMA", 0xB0000), ("doubleMMA", 0x80000)):
    _lib.set_option("debug_flags", dbg | 0x40000)
    for _ in range(5): s = cb.maxsim(qb, bank)
    torch.cuda.synchronize()
    f = s.flatten().double(); tot = f[:296].view(148, 2); x = f[512:512 + 8 * 148].view(148, 8)
    jobs = 28 * 8
    print(f"{name:24s}: CTA cycles max {tot[:,0].max():.0f} -> {tot[:,0].max()/jobs:.0f} cycles/job (224 jobs on the longest CTA)", flush=True)
_lib.set_option("debug_flags", 0)
