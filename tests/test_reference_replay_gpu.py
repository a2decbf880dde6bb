"""The UNMODIFIED reference package (baseline/_ref/colpali_engine, put there by baseline/install_ref.py; it travels to the
GPU box) patched by ``colpali_b200.install()``, then the reference's OWN offline hot-path tests replayed through the
patched attributes:

* tests/utils/test_processing_utils.py:8-35  (scorer shapes, list == tensor equivalence),
* tests/loss/test_li_losses.py:75-181        (loss KATs: zero embeddings -> ln B / softplus(0), with/without filtering,
  explicit negatives with and without the in-batch term),
* tests/loss/test_bi_losses.py:42-131        (bi-encoder loss KATs, row f-4).

The reference tests build CPU tensors; the B200 losses have no CPU path, so the loss tests run under
``torch.device("cuda")`` as the default device (their tensor factories then allocate on the GPU, nothing else changes).
``TestColbertModule`` exercises the reference's eager helper methods on the base class, which ``install()`` leaves
untouched; it is replayed on the CPU to show the patch does not disturb it.
"""
import importlib.util
import inspect
import os
import sys

import pytest
import torch

import colpali_b200 as cb
from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")


def _load(name):
    spec = importlib.util.spec_from_file_location(f"ref_replay_{name}", os.path.join(REF, "ref_tests", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def patched_reference():
    if not os.path.isdir(os.path.join(REF, "colpali_engine")):
        pytest.skip("baseline/_ref is not installed (python baseline/install_ref.py)")
    sys.path.insert(0, REF)
    import colpali_engine  # noqa: F401  -- the real package, not a stub
    import colpali_engine.loss as L
    from colpali_engine.utils.processing_utils import BaseVisualRetrieverProcessor as P

    assert os.path.realpath(colpali_engine.__file__).startswith(os.path.realpath(REF))
    orig_scorer = P.score_multi_vector
    cb.install()
    try:
        assert P.score_multi_vector is cb.score_multi_vector and P.score_multi_vector is not orig_scorer
        assert L.ColbertLoss is cb.ColbertLoss and L.ColbertPairwiseNegativeCELoss is cb.ColbertPairwiseNegativeCELoss
        yield
    finally:
        cb.uninstall()
        sys.path.remove(REF)
    assert P.score_multi_vector is orig_scorer


def test_reference_scorer_tests_through_the_patch(patched_reference):
    from colpali_b200 import _lib

    before = _lib.gpu_launches()
    mod = _load("test_processing_utils")
    torch.manual_seed(0)
    mod.test_score_multi_vector_embeddings()   # device=None -> cuda:0, returns CPU fp32 (processing_utils.py:161,180)
    mod.test_score_single_vector_embeddings()  # patched too: the dense fp32 kernel (any dim, operands keep their dtype)
    assert _lib.gpu_launches() - before == 3   # two score_multi_vector calls + one score_single_vector call


def test_reference_loss_kats_through_the_patch(patched_reference):
    from colpali_b200 import _lib

    mod = _load("test_li_losses")
    assert mod.ColbertLoss is cb.ColbertLoss  # the test module imported the patched names
    before = _lib.gpu_launches()
    ran = 0
    with torch.device("cuda"):
        for cname in ("TestColbertLoss", "TestColbertNegativeCELoss", "TestColbertPairwiseCELoss",
                      "TestColbertPairwiseNegativeCELoss"):
            inst = getattr(mod, cname)()
            for name, fn in inspect.getmembers(inst, inspect.ismethod):
                if name.startswith("test_"):
                    fn()
                    ran += 1
    assert ran == 7
    assert _lib.gpu_launches() > before
    inst = mod.TestColbertModule()  # the reference's own base class and helpers, untouched by install()
    for name, fn in inspect.getmembers(inst, inspect.ismethod):
        if name.startswith("test_"):
            fn()


def test_reference_bi_loss_kats_through_the_patch(patched_reference):
    from colpali_b200 import _lib

    mod = _load("test_bi_losses")
    assert mod.BiEncoderLoss is cb.BiEncoderLoss and mod.BiPairwiseNegativeCELoss is cb.BiPairwiseNegativeCELoss
    before = _lib.gpu_launches()
    ran = 0
    with torch.device("cuda"):
        for cname in ("TestBiEncoderLoss", "TestBiNegativeCELoss", "TestBiPairwiseCELoss", "TestBiPairwiseNegativeCELoss"):
            inst = getattr(mod, cname)()
            for name, fn in inspect.getmembers(inst, inspect.ismethod):
                if name.startswith("test_"):
                    fn()
                    ran += 1
    assert ran == 8
    assert _lib.gpu_launches() > before
    inst = mod.TestBiEncoderModule()  # the reference's own base class and helpers, untouched by install()
    for name, fn in inspect.getmembers(inst, inspect.ismethod):
        if name.startswith("test_"):
            fn()


def test_reference_scorer_equals_patched_scorer_on_cfg1(patched_reference):
    """Same call, reference vs patched, on BASELINE configs[0] (fp32 inputs to the reference = the fp32 oracle)."""
    from colpali_engine.utils.processing_utils import BaseVisualRetrieverProcessor as P
    from oracle import li_oracle as O

    q, d = O.cfg1_inputs()
    got = P.score_multi_vector(q, d)
    cb.uninstall()
    try:
        want = P.score_multi_vector(q.float(), d.float(), device="cpu")
    finally:
        cb.install()
    assert got.shape == want.shape and got.dtype == want.dtype == torch.float32 and got.device == want.device
    assert ((got - want).abs() / want.abs()).max() < 1e-5
    assert torch.equal(got.argmax(1), want.argmax(1))
