"""Per-job cycle budget of the MaxSim kernel at cfg2 from its in-kernel %clock64 counters (debug flag 0x40000: CTA b writes
its counters into the score buffer instead of scores).  Cycles do not depend on the power-limited clock, so this is the
A/B that timing cannot give.  1024 cycles = 8 MMAs of 128 = the ideal job.  One JSON line per boundary mode."""
import json, sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from colpali_b200 import _lib
from oracle import li_oracle as O

dev = torch.device("cuda:0")
q, d = O.cfg2_inputs()
qb, bank = cb.QueryBlock(q.to(dev), dev), cb.DocBank.from_passages(d.to(dev), dev)
for n_d, docs in ((1030, d), (1024, d[:, :1024].contiguous())):
    bk = cb.DocBank.from_passages(docs.to(dev), dev)
    for bmode in (0, 1):
        _lib.set_option("boundary_mode", bmode)
        _lib.set_option("debug_flags", 0x40000)
        try:
            raw = cb.maxsim(qb, bk)
            torch.cuda.synchronize()
        finally:
            _lib.set_option("debug_flags", 0)
        raw = raw.flatten().cpu()
        n_cta = 148
        cyc = raw[0:2 * n_cta:2]; ns = raw[1:2 * n_cta:2]
        c = raw[512:512 + 8 * n_cta].view(n_cta, 8)   # w_full, w_tmem, e_wait, e_hold, e_post, e_hold2, n_path2, jobs
        jobs = c[:, 7].clamp_min(1)
        print(json.dumps({
            "n_d": n_d, "boundary_mode": bmode, "ctas": int((c[:, 7] > 0).sum()),
            "cycles_per_job_mean": float((cyc / jobs)[c[:, 7] > 0].mean()), "cycles_per_cta_max": float(cyc.max()),
            "ghz_in_kernel": float((cyc / ns.clamp_min(1)).median()),
            "issuer_wait_full_per_job": float((c[:, 0] / jobs).mean()), "issuer_wait_tmem_per_job": float((c[:, 1] / jobs).mean()),
            "epilogue_wait_per_job": float((c[:, 2] / jobs).mean()),
            "hold_plain_per_job": float((c[:, 3] / (jobs - c[:, 6]).clamp_min(1)).mean()),
            "hold_boundary_per_job": float((c[:, 5] / c[:, 6].clamp_min(1)).mean()), "boundary_jobs_share": float((c[:, 6] / jobs).mean()),
        }), flush=True)
_lib.set_option("boundary_mode", 1)

# ---- the CTA-pair kernel (option pair=1) at cfg2: leader and peer CTAs apart ------------------------------------------
_lib.set_option("pair", 1)
_lib.set_option("debug_flags", 0x40000)
try:
    raw = cb.maxsim(qb, bank)
    torch.cuda.synchronize()
finally:
    _lib.set_option("debug_flags", 0)
    _lib.set_option("pair", 0)
raw = raw.flatten().cpu()
n_cta = 148
cyc = raw[0:2 * n_cta:2]
c = raw[512:512 + 8 * n_cta].view(n_cta, 8)
for name, sel in (("leader", slice(0, n_cta, 2)), ("peer", slice(1, n_cta, 2))):
    cc, cy = c[sel], cyc[sel]
    jobs = cc[:, 7].clamp_min(1)
    print(json.dumps({"mode": "pair, cfg2", "cta": name, "cycles_per_job_mean": float((cy / jobs).mean()),
                      "issuer_wait_full_per_job": float((cc[:, 0] / jobs).mean()), "issuer_wait_tmem_per_job": float((cc[:, 1] / jobs).mean()),
                      "epilogue_wait_per_job": float((cc[:, 2] / jobs).mean()),
                      "hold_plain_per_job": float((cc[:, 3] / (jobs - cc[:, 6]).clamp_min(1)).mean()),
                      "hold_boundary_per_job": float((cc[:, 5] / cc[:, 6].clamp_min(1)).mean())}), flush=True)

# ---- the training forward (argmax mode) at cfg3: 64 queries x 64 left-padded documents; counters of warp 2 = epilogue
# group 0, which folds query tile 0 of every document tile (half of the CTA's jobs) --------------------------------------
q3, d3, _ = O.cfg3_inputs()
qb3, bank3 = cb.QueryBlock(q3.to(dev), dev), cb.DocBank.from_passages(d3.to(dev), dev)
_lib.set_option("debug_flags", 0x40000)
try:
    raw, _am = cb.maxsim(qb3, bank3, want_argmax=True)
    torch.cuda.synchronize()
finally:
    _lib.set_option("debug_flags", 0)
raw = raw.flatten().cpu()
n_cta = 144
cyc = raw[0:2 * n_cta:2]
ns = raw[1:2 * n_cta:2]
c = raw[512:512 + 8 * n_cta].view(n_cta, 8)
own = c[:, 7].clamp_min(1)
for _ in range(5): cb.maxsim(qb3, bank3, want_argmax=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): cb.maxsim(qb3, bank3, want_argmax=True)
e1.record(); torch.cuda.synchronize()
warm_us = 1e3 * e0.elapsed_time(e1) / 50
print(json.dumps({"mode": "argmax, cfg3", "kernel_us_back_to_back": warm_us, "cta_main_loop_us_mean": float(ns.mean()) / 1e3,
                  "cta_main_loop_us_max": float(ns.max()) / 1e3, "ctas": int((c[:, 7] > 0).sum()), "own_jobs_per_warp": float(own.mean()),
                  "cta_cycles_mean": float(cyc.mean()), "cta_cycles_max": float(cyc.max()),
                  "cycles_per_own_job": float((cyc / own).mean()),
                  "hold_whole_tile_per_job": float((c[:, 3] / (own - c[:, 6]).clamp_min(1)).mean()),
                  "hold_other_per_job": float((c[:, 5] / c[:, 6].clamp_min(1)).mean()), "other_jobs_share": float((c[:, 6] / own).mean()),
                  "wait_for_mma_per_own_job": float((c[:, 2] / own).mean()),
                  "post_release_per_own_job": float((c[:, 4] / own).mean())}), flush=True)
