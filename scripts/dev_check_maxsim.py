"""Developer check (GPU): fused MaxSim vs a plain fp32 torch computation on several shapes + timing."""
import sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import colpali_b200 as cb

dev = torch.device("cuda:0")

def ref_scores(qs, ps):  # lists of [len, D] bf16 -> fp32 scores (no padding semantics)
    out = torch.empty(len(qs), len(ps))
    for i, q in enumerate(qs):
        for j, p in enumerate(ps):
            out[i, j] = (q.float() @ p.float().T).amax(1).sum()
    return out

def check(name, nq_list, nd_list, D=128, seed=0):
    g = torch.Generator().manual_seed(seed)
    qs = [F.normalize(torch.randn(n, D, generator=g), dim=-1).bfloat16() for n in nq_list]
    ps = [F.normalize(torch.randn(n, D, generator=g), dim=-1).bfloat16() for n in nd_list]
    bank = cb.DocBank.from_passages([p.to(dev) for p in ps], dev, reference_padding=False)
    q = cb.QueryBlock([x.to(dev) for x in qs], dev)
    s, am = cb.maxsim(q, bank, want_argmax=True)
    s2 = cb.maxsim(q, bank)
    torch.cuda.synchronize()
    r = ref_scores([x.to(dev) for x in qs], [p.to(dev) for p in ps])
    e1 = (s.cpu() - r).abs().max().item(); e2 = (s2.cpu() - r).abs().max().item()
    print(f"{name}: max abs err argmax-variant {e1:.3e} plain {e2:.3e} (scores ~{r.abs().mean():.2f})", flush=True)
    return e1 < 1e-3 and e2 < 1e-3

ok = True
ok &= check("cfg1 4x16 Nq32 Nd256", [32]*4, [256]*16)
ok &= check("ragged small", [5, 32, 17, 1], [1, 16, 17, 255, 256, 257, 600, 1030, 7])
ok &= check("nq>32", [40, 70, 33], [300, 511, 513])
ok &= check("many queries", [32]*9, [100]*50)
ok &= check("D=32", [2, 4], [8, 4, 16], D=32)
print("ALL OK" if ok else "MISMATCH", flush=True)

# timing at cfg2
torch.manual_seed(0)
q = F.normalize(torch.randn(32, 32, 128, device=dev), dim=-1).bfloat16()
d = F.normalize(torch.randn(1000, 1030, 128, device=dev), dim=-1).bfloat16()
bank = cb.DocBank.from_passages(d, dev)
qb = cb.QueryBlock(q, dev)
for _ in range(3): s = cb.maxsim(qb, bank)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): s = cb.maxsim(qb, bank)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fl = 2 * 32 * 32 * 1000 * 1030 * 128
print(f"cfg2 kernel {ms:.4f} ms  -> {fl/ms/1e9:.1f} TFLOP/s, {32/ms*1e3:.0f} queries/s", flush=True)
r = torch.einsum("bnd,csd->bcns", q.float(), d[:50].float()).amax(3).sum(2)
print("cfg2 err vs fp32 (first 50 docs):", (s[:, :50] - r).abs().max().item())
