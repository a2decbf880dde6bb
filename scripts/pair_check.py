"""CTA-pair MaxSim kernel (option pair=1) against the single-CTA kernel and fp32 torch, then interleaved timings.

Usage: python scripts/pair_check.py [check|perf]   (run under `timeout`: a protocol bug traps after 4 s)"""
import json, statistics, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
from oracle import li_oracle as O

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "check"


def fp32(q, d):  # [Q, Nq, D] x [C, Nd, D] -> [Q, C]
    return torch.einsum("bnd,csd->bcns", q.float(), d.float()).amax(3).sum(2)


def one(n_q, n_d, nq_tok, nd_tok, seed, ragged=False):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n_q, nq_tok, 128, generator=g).bfloat16().to(dev)
    if ragged:
        lens = torch.randint(max(1, nd_tok // 2), nd_tok + 1, (n_d,), generator=g).tolist()
        docs = [torch.randn(l, 128, generator=g).bfloat16().to(dev) for l in lens]
        padded = torch.nn.utils.rnn.pad_sequence(docs, batch_first=True)
        want = torch.stack([torch.einsum("bnd,sd->bns", q.float(), x.float()).amax(2).sum(1) for x in docs], 1)
    else:
        docs = torch.randn(n_d, nd_tok, 128, generator=g).bfloat16().to(dev)
        want = fp32(q, docs)
    qb, bank = cb.QueryBlock(q, dev), cb.DocBank.from_passages(docs, dev)
    _lib.set_option("pair", 0)
    a = cb.maxsim(qb, bank).clone()
    _lib.set_option("pair", 1)
    b = cb.maxsim(qb, bank).clone()
    torch.cuda.synchronize()
    err = (b - want).abs().max().item() / max(1.0, want.abs().max().item())
    ok = torch.equal(a, b) and err < 2e-6
    print(json.dumps({"case": [n_q, n_d, nq_tok, nd_tok, ragged], "bit_equal_to_single": torch.equal(a, b),
                      "max_abs_diff_single": (a - b).abs().max().item(), "rel_err_fp32": err, "ok": ok}), flush=True)
    return ok


if what == "check":
    ok = True
    ok &= one(8, 16, 32, 256, 1)          # 2 query tiles (R = 1 per CTA), whole tiles
    ok &= one(16, 40, 32, 1030, 2)        # 4 query tiles (R = 2), boundaries, balanced partitions
    ok &= one(32, 200, 32, 1030, 3)       # 8 query tiles
    ok &= one(16, 64, 32, 700, 4, True)   # ragged lengths
    ok &= one(128, 300, 32, 1030, 5)      # cfg4 geometry (32 query tiles)
    q, d = O.cfg2_inputs()
    qb, bank = cb.QueryBlock(q.to(dev), dev), cb.DocBank.from_passages(d.to(dev), dev)
    _lib.set_option("pair", 0); a = cb.maxsim(qb, bank).clone()
    _lib.set_option("pair", 1); b = cb.maxsim(qb, bank).clone()
    torch.cuda.synchronize()
    print(json.dumps({"case": "cfg2", "bit_equal_to_single": torch.equal(a, b)}), flush=True)
    ok &= torch.equal(a, b)
    print("PAIR CHECK", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)

q, d = O.cfg2_inputs()
qb, bank = cb.QueryBlock(q.to(dev), dev), cb.DocBank.from_passages(d.to(dev), dev)
FLOPS = 2.0 * 32 * 32 * 1000 * 1030 * 128
BURST, ROUNDS = 25, 24
VARIANTS = [dict(pair=pr, early_spin=sp) for pr in (0, 1) for sp in (1 << 30, 512, 128, 0)]
times = [[] for _ in VARIANTS]
for r in range(ROUNDS):
    for i, v in enumerate(VARIANTS):
        for k, x in v.items(): _lib.set_option(k, x)
        for _ in range(3): cb.maxsim(qb, bank, independent=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(BURST): cb.maxsim(qb, bank, independent=True)
        e1.record(); torch.cuda.synchronize()
        times[i].append(e0.elapsed_time(e1) / BURST)
for i, v in enumerate(VARIANTS):
    t = times[i]
    print(json.dumps({**v, "median_ms": round(statistics.median(t), 5), "min_ms": round(min(t), 5),
                      "median_pflops": round(FLOPS / statistics.median(t) / 1e12, 4),
                      "median_ratio_to_first": round(statistics.median(a / b for a, b in zip(t, times[0])), 4)}), flush=True)
_lib.set_option("pair", 0); _lib.set_option("early_spin", 0)
