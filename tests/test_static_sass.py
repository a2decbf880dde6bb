"""CPU: what the compiler made of the kernels, read from the built library with cuobjdump (no GPU).

Pins the claims DESIGN.md makes about the hot path: the MaxSim and head kernels are tcgen05 kernels (``UTCHMMA`` =
tcgen05.mma, ``LDTM`` = tcgen05.ld from TMEM, ``UTMALDG`` = TMA tensor loads, ``SYNCS`` = mbarriers -- the mnemonics
B200_PROFILING.md lists), nothing on the product path is a warp-level ``mma.sync`` kernel except the two smooth-max
recompute kernels DESIGN 4.3 names, and the headline kernel ``maxsim_fwd_kernel<2, 0>`` neither spills nor keeps a
stack frame.  ``profiles/r02_static_sass.md`` is the same data as a table (scripts/static_report.py).
"""
import importlib.util
import os
import re
import shutil

import pytest

from conftest import ROOT
from colpali_b200 import build as cbuild


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    cbuild.build(force=False)
    spec = importlib.util.spec_from_file_location("static_report", os.path.join(ROOT, "scripts", "static_report.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    counts, total = mod.sass_counts()
    res = mod.resources()
    names = list(counts)
    pretty = dict(zip(names, (mod.short(n) for n in mod.demangle(names))))
    return {pretty[n]: (counts[n], total[n], res.get(n)) for n in names}


def _family(sass, pattern):
    hit = {k: v for k, v in sass.items() if re.search(pattern, k)}
    assert hit, f"no kernel matches {pattern!r} in {sorted(sass)}"
    return hit


@pytest.mark.parametrize("pattern", [r"cpb::maxsim_fwd_kernel<", r"kpipe::maxsim_kpipe_kernel<", r"pair::maxsim_pair_kernel<",
                                     r"cpb::head_fwd_kernel$", r"cpb::head_wide_kernel$"])
def test_tensor_kernels_are_tcgen05_tma_tmem_kernels(sass, pattern):
    for name, (c, _, _) in _family(sass, pattern).items():
        assert c["UTCHMMA"] > 0, f"{name}: no tcgen05.mma"
        assert c["UTMALDG"] > 0, f"{name}: no TMA tensor load"
        assert c["LDTM"] > 0, f"{name}: accumulators are not read from TMEM"
        assert c["UTCBAR"] > 0 and c["SYNCS"] > 0, f"{name}: no tcgen05.commit / mbarrier pipeline"
        assert c["HMMA"] == 0, f"{name}: warp-level mma.sync in a tcgen05 kernel"


def test_all_mode_and_shape_variants_are_built(sass):
    assert len(_family(sass, r"cpb::maxsim_fwd_kernel<")) == 6      # R in {1, 2} x {max, argmax, smooth}
    assert len(_family(sass, r"kpipe::maxsim_kpipe_kernel<")) == 9  # K panels {3, 4, 5} = dims 192 / 256 / 320 x 3 modes
    assert len(_family(sass, r"pair::maxsim_pair_kernel<")) == 6
    assert len(_family(sass, r"dense_tile_kernel<")) == 8           # {128, 64}-tiles x {fp32, bf16}^2 operands


def test_only_the_smooth_backward_uses_mma_sync(sass):
    hmma = sorted(k for k, (c, _, _) in sass.items() if c["HMMA"])
    assert hmma == ["cpb::smooth_bwd_dd_kernel", "cpb::smooth_bwd_dq_kernel"], hmma


def test_headline_kernel_has_no_spills(sass):
    c, n_instr, res = sass["cpb::maxsim_fwd_kernel<2, 0>"]  # cfg2: two resident query tiles, plain max
    regs, stack, _, _ = res
    assert stack == 0 and c["LDL"] == 0 and c["STL"] == 0
    assert regs <= 255 and n_instr > 1000
    # the smooth-max variants are the ones with the exponentials (MUFU.EX2), the plain-max kernel keeps only the loss tail's
    assert sass["cpb::maxsim_fwd_kernel<2, 2>"][0]["MUFU.EX2"] > c["MUFU.EX2"]


def test_multi_gpu_epilogue_and_exchange_use_reductions_not_cas_loops(sass):
    # completion counters / peer scatter: red.global (REDG), never a compare-and-swap loop
    for name in ("cpb::exchange_push_kernel", "cpb::signal_peers_kernel"):
        c = sass[name][0]
        assert c["REDG"] > 0 and c["ATOMG"] == 0, name
