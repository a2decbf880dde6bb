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
