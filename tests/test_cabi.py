"""CPU: the C-ABI library builds, loads without a GPU driver, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from colpali_b200 import _lib
from colpali_b200 import build as cbuild


@pytest.fixture(scope="module")
def lib():
    cbuild.build(force=False)
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "colpali_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpb_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib):
    names = header_functions()
    assert names, "no functions parsed from include/colpali_b200.h"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"libcolpali_b200.so does not export {n}"


def test_abi_version(lib):
    assert lib.cpb_abi_version() == 1


def test_argument_validation_needs_no_gpu(lib):
    # invalid shapes are rejected before any CUDA call
    rc = lib.cpb_maxsim_fwd(None, 0, 32, None, 0, None, None, None, 0, None, None, None, 0, None)
    assert rc == -1 and b"positive" in lib.cpb_last_error()
    rc = lib.cpb_maxsim_fwd(None, 1, 33, None, 0, None, None, None, 1, None, None, None, 0, None)
    assert rc == -1 and b"multiple of 32" in lib.cpb_last_error()
    assert lib.cpb_maxsim_workspace_bytes(4, 32, 16) == 0
    assert lib.cpb_maxsim_workspace_bytes(4, 96, 16) == 3 * 4 * 16 * 4


def test_product_path_fails_loudly_on_cpu():
    import torch

    import colpali_b200 as cb

    q = [torch.randn(4, 128)]
    with pytest.raises(cb.ColpaliB200Error):
        cb.score_multi_vector(q, q, device="cpu")
    with pytest.raises(ValueError, match="No queries"):
        cb.score_multi_vector([], q, device="cpu")
    with pytest.raises(ValueError, match="No passages"):
        cb.score_multi_vector(q, [], device="cpu")


def test_every_entry_point_validates_before_touching_cuda(lib):
    """Bad arguments are rejected with CPB_E_INVALID / CPB_E_UNSUPPORTED and a message, without a GPU."""
    INVALID, UNSUPPORTED = -1, -2
    # losses
    rc = lib.cpb_colbert_loss_fwd(None, None, 4, 32, 3, 0, 0.02, 1, 0, 0.95, 0.5, 0, None, None, None, None)
    assert rc == INVALID and b"positive index out of range" in lib.cpb_last_error()          # offset + B > C
    rc = lib.cpb_colbert_loss_fwd(None, None, 4, 32, 4, 7, 0.02, 1, 0, 0.95, 0.5, 0, None, None, None, None)
    assert rc == INVALID and b"unknown loss mode" in lib.cpb_last_error()
    rc = lib.cpb_colbert_loss_fwd(None, None, 4, 32, 6, 2, 0.02, 1, 0, 0.95, 0.5, 0, None, None, None, None)
    assert rc == INVALID and b"sigmoid" in lib.cpb_last_error()                               # needs a square matrix
    rc = lib.cpb_colbert_loss_fwd(None, None, 4, 32, 4, 0, -1.0, 1, 0, 0.95, 0.5, 0, None, None, None, None)
    assert rc == INVALID and b"temperature" in lib.cpb_last_error()
    rc = lib.cpb_colbert_neg_loss_fwd(None, None, None, 4, 32, 4, 2, 0, 0.02, 1, 0, 0.95, 0.5, 0.5, 0, None, None, None, None)
    assert rc == INVALID
    # backward
    rc = lib.cpb_maxsim_bwd(None, None, None, None, 4, 32, None, 10, None, 3, None, None, None)
    assert rc == INVALID and b"null device pointer" in lib.cpb_last_error()
    # head
    rc = lib.cpb_head_fwd(None, 0, 1536, None, None, 128, None, None, None, 0, None)
    assert rc == INVALID
    rc = lib.cpb_head_fwd(None, 10, 1536, None, None, 352, None, None, None, 0, None)
    assert rc == UNSUPPORTED and b"352" in lib.cpb_last_error()                               # above ColQwen3's 320
    rc = lib.cpb_head_fwd(None, 10, 1536, None, None, 200, None, None, None, 0, None)
    assert rc == UNSUPPORTED and b"200" in lib.cpb_last_error()                               # not a multiple of 32
    rc = lib.cpb_head_fwd(None, 10, 1000, None, None, 128, None, None, None, 0, None)
    assert rc == UNSUPPORTED and b"multiple of 64" in lib.cpb_last_error()
    # balanced / all-gather variants and their helpers
    rc = lib.cpb_maxsim_fwd_balanced(None, 1, 32, None, 0, None, None, None, 1, None, None, None, 0, 0, 0, None, 0, 0, None)
    assert rc == INVALID and b"epoch" in lib.cpb_last_error()
    rc = lib.cpb_maxsim_fwd_allgather(None, 1, 32, None, 0, None, None, None, 1, None, 2, 0, 0, 0, 0, None, 0, 1, None, 0, 1, None)
    assert rc == INVALID and b"peer" in lib.cpb_last_error()
    assert lib.cpb_wait_flags(None, 2, 1, None) == INVALID
    assert lib.cpb_maxsim_split_workspace_bytes(32, 32) > 0
    # tuning knobs
    assert lib.cpb_set_option(b"cluster", 3) == INVALID and lib.cpb_set_option(b"no_such_option", 1) == INVALID
    assert lib.cpb_set_option(b"cluster", 0) == 0 and lib.cpb_set_option(b"balanced", 1) == 0
