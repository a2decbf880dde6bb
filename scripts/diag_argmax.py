"""Diagnosis: argmax of the training forward at cfg3 against torch, mismatches classified by position."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
from oracle import li_oracle as O

dev = torch.device("cuda:0")
q, d, lens = O.cfg3_inputs()
q, d = q.to(dev), d.to(dev)
qb, bank = cb.QueryBlock(q, dev), cb.DocBank.from_passages(d, dev)
sim = torch.einsum("rd,csd->crs", q.float().reshape(-1, 128), d.float())   # [C, rows, L]
want = sim.argmax(2).int()
wmax = sim.amax(2)
for name, opts in (("default", {}), ("unbalanced", {"balanced": 0}), ("cluster1", {"cluster": 1})):
    for k, v in opts.items(): _lib.set_option(k, v)
    s, am = cb.maxsim(qb, bank, want_argmax=True)
    torch.cuda.synchronize()
    for k in opts: _lib.set_option(k, 1 if k == "balanced" else 0)
    bad = (am != want)
    out = {"variant": name, "n_bad": int(bad.sum()), "n": bad.numel()}
    if bad.any():
        c, r = bad.nonzero(as_tuple=True)
        got_i, want_i = am[c, r].long(), want[c, r].long()
        gv = sim[c, r, got_i.clamp_min(0)]
        out["value_at_got_equals_max"] = float((gv == wmax[c, r]).float().mean())
        out["got_minus_want_hist"] = torch.unique((got_i - want_i), return_counts=True)[0][:20].tolist()
        out["want_mod32"] = torch.bincount((want_i + 0) % 32, minlength=32).tolist()
        out["want_abs_mod256"] = torch.bincount((c * 1030 + want_i) % 256 // 32, minlength=8).tolist()
        out["got_neg"] = int((got_i < 0).sum())
        out["examples"] = [(int(c[i]), int(r[i]), int(got_i[i]), int(want_i[i]), float(gv[i]), float(wmax[c[i], r[i]])) for i in range(min(12, c.numel()))]
        out["bad_docs"] = torch.unique(c).tolist()[:40]
    print(json.dumps(out), flush=True)
