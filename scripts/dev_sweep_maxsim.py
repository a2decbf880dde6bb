"""Developer sweep (GPU): cfg2 kernel time under tuning knobs (cluster size, query tiles per CTA, epilogue off)."""
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

for nd in (1030, 1024):
    d = F.normalize(torch.randn(1000, nd, 128, device=dev), dim=-1).bfloat16()
    bank = cb.DocBank.from_passages(d, dev)
    qb = cb.QueryBlock(q, dev)
    ref = torch.einsum("bnd,csd->bcns", q.float(), d[:40].float()).amax(3).sum(2)
    fl = 2 * 32 * 32 * 1000 * nd * 128
    for cluster in (1, 2, 4):
        for r in (1, 2):
            for dbg in (0, 0x10000):
                _lib.set_option("cluster", cluster); _lib.set_option("qtiles_per_cta", r); _lib.set_option("debug_flags", dbg)
                try:
                    ms, s = timeit(qb, bank)
                    err = (s[:, :40] - ref).abs().max().item() if not dbg else float("nan")
                    print(f"Nd={nd} cluster={cluster} R={r} skip_epi={bool(dbg)}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s  err={err:.2e}", flush=True)
                except Exception as e:
                    print(f"Nd={nd} cluster={cluster} R={r} dbg={dbg}: FAILED {e}", flush=True)
_lib.set_option("cluster", 0); _lib.set_option("qtiles_per_cta", 0); _lib.set_option("debug_flags", 0)
# bigger query batches: 128 queries (32 query tiles)
q128 = F.normalize(torch.randn(128, 32, 128, device=dev), dim=-1).bfloat16()
d = F.normalize(torch.randn(1000, 1030, 128, device=dev), dim=-1).bfloat16()
bank = cb.DocBank.from_passages(d, dev)
for cluster in (1, 2, 4):
    _lib.set_option("cluster", cluster)
    ms, s = timeit(cb.QueryBlock(q128, dev), bank, 20)
    print(f"128 queries cluster={cluster}: {ms*1e3:.1f} us {2*128*32*1000*1030*128/ms/1e9:.0f} TFLOP/s", flush=True)
