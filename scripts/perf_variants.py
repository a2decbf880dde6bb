"""A/B timings of the MaxSim kernel's launch / epilogue options at cfg2 (32 q x 1000 docs x 1030 x 128).

The SM clock drifts by 10 % within a second under the power governor, so variants are INTERLEAVED: every round times a
short burst of each variant; reported are the per-variant medians and the median of the per-round ratios to the first
variant.  One JSON line per variant."""
import json, statistics, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
from oracle import li_oracle as O

dev = torch.device("cuda:0")
q, d = O.cfg2_inputs()
qb, bank = cb.QueryBlock(q.to(dev), dev), cb.DocBank.from_passages(d.to(dev), dev)
FLOPS = 2.0 * 32 * 32 * 1000 * 1030 * 128
VARIANTS = [("pdl0_b0", 0, False, 0), ("pdl0_b1", 0, False, 1), ("pdl1_b1", 1, False, 1), ("pdl1_indep_b1", 1, True, 1),
            ("pdl1_indep_b0", 1, True, 0)]
BURST, ROUNDS = 25, 40


def burst(pdl, indep, bmode):
    _lib.set_option("pdl", pdl); _lib.set_option("boundary_mode", bmode)
    for _ in range(3): cb.maxsim(qb, bank, independent=indep)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(BURST): out = cb.maxsim(qb, bank, independent=indep)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / BURST, out


times = {v[0]: [] for v in VARIANTS}
ref = None
for r in range(ROUNDS):
    for name, pdl, indep, bmode in VARIANTS:
        ms, out = burst(pdl, indep, bmode)
        if ref is None: ref = out.clone()
        assert torch.equal(out, ref), name
        times[name].append(ms)
base = times[VARIANTS[0][0]]
for name, *_ in VARIANTS:
    t = times[name]
    print(json.dumps({"variant": name, "median_ms": statistics.median(t), "min_ms": min(t),
                      "median_tflops": FLOPS / statistics.median(t) / 1e9,
                      "median_ratio_to_first": statistics.median(a / b for a, b in zip(t, base))}), flush=True)
_lib.set_option("pdl", 1); _lib.set_option("boundary_mode", 1)
