"""A/B timings of the MaxSim kernel's launch/epilogue options at cfg2 (32 q x 1000 docs x 1030 x 128), same process, same
clocks: programmatic dependent launch off / on / independent, boundary-tile path 0 / 1.  One JSON line per variant."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
from oracle import li_oracle as O

dev = torch.device("cuda:0")
q, d = O.cfg2_inputs()
qb, bank = cb.QueryBlock(q.to(dev), dev), cb.DocBank.from_passages(d.to(dev), dev)
FLOPS = 2.0 * 32 * 32 * 1000 * 1030 * 128
base = None


def timed(independent, n=200, warm=10):
    for _ in range(warm): cb.maxsim(qb, bank, independent=independent)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = cb.maxsim(qb, bank, independent=independent)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ref = None
for rep in range(2):  # twice: the second pass shows how much is clock drift
    for pdl, indep, bmode in ((0, False, 0), (0, False, 1), (1, False, 1), (1, True, 1), (1, True, 0)):
        _lib.set_option("pdl", pdl); _lib.set_option("boundary_mode", bmode)
        ms, out = timed(indep)
        if ref is None: ref = out.clone()
        print(json.dumps({"rep": rep, "pdl": pdl, "independent": indep, "boundary_mode": bmode, "ms": ms,
                          "tflops": FLOPS / ms / 1e9, "bit_equal_to_first": bool(torch.equal(out, ref))}), flush=True)
_lib.set_option("pdl", 1); _lib.set_option("boundary_mode", 1)
