"""Developer probe (GPU): sweep the issuer's split point (cpb_set_option mma_split) at cfg2 shapes."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
dev = torch.device("cuda:0")
torch.manual_seed(0)
q = F.normalize(torch.randn(32, 32, 128, device=dev), dim=-1).bfloat16()
for nd in (1024, 1030):
    d = F.normalize(torch.randn(1000, nd, 128, device=dev), dim=-1).bfloat16()
    bank = cb.DocBank.from_passages(d, dev); qb = cb.QueryBlock(q, dev)
    ref = torch.einsum("bnd,csd->bcns", q.float(), d[:20].float()).amax(3).sum(2)
    for split in (5, 6, 7, 8):
      _lib.set_option("mma_split", split)
      for name, dbg in (("normal", 0), ("noEpi", 0x10000)):
        _lib.set_option("debug_flags", 0x40000 | dbg)
        for _ in range(5): s = cb.maxsim(qb, bank)
        torch.cuda.synchronize()
        f = s.flatten().double()
        tot = f[:296].view(148, 2); x = f[512:512 + 8 * 148].view(148, 8); jobs = x[:, 7]; n2 = x[:, 6]
        _lib.set_option("debug_flags", dbg)
        s2 = cb.maxsim(qb, bank); torch.cuda.synchronize()
        err = (s2[:, :20] - ref).abs().max().item() if not dbg else float("nan")
        print(f"Nd={nd} split={split} {name:7s}: cycles/job {(tot[:,0]/jobs).mean():6.0f} (max CTA {tot[:,0].max():.0f}) | issuer blocked/job: TMA {(x[:,0]/jobs).mean():4.0f} epi {(x[:,1]/jobs).mean():4.0f} | "
              f"epilogue/job: wait {(x[:,2]/jobs).mean():4.0f} hold {(x[:,3]/(jobs-n2)).mean():4.0f} hold2 {(x[:,5]/n2.clamp_min(1)).mean():4.0f} post {(x[:,4]/jobs).mean():4.0f} err {err:.1e}", flush=True)
_lib.set_option("mma_split", 6); _lib.set_option("debug_flags", 0)
