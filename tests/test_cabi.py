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
    assert lib.cpb_abi_version() == 3


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
    import ctypes

    from colpali_b200._lib import DenseDotArgs, LossDesc, MaxSimArgs, MaxSimBwdArgs

    INVALID, UNSUPPORTED = -1, -2
    err = lib.cpb_last_error
    # losses (struct argument blocks: struct_size is checked first)
    d = LossDesc(mode=0, temperature=0.02, normalize_scores=1, filter_threshold=0.95, filter_factor=0.5)

    def loss(n_q, n_docs, dim=128):
        return lib.cpb_colbert_loss_launch(ctypes.byref(d), None, 16, n_q, 32, n_docs, dim, None)  # d_q non-null, never read

    def bi_loss(n_q, n_docs):  # d_q == NULL: scores of single vectors (bi-encoder losses)
        return lib.cpb_colbert_loss_launch(ctypes.byref(d), None, None, n_q, 0, n_docs, 0, None)

    assert loss(4, 3) == INVALID and b"positive index out of range" in err()                    # offset + B > C
    d.mode = 7
    assert loss(4, 4) == INVALID and b"unknown loss mode" in err()
    d.mode = 2
    assert loss(4, 6) == INVALID and b"sigmoid" in err()                                        # needs a square matrix
    d.mode, d.temperature = 0, -1.0
    assert loss(4, 4) == INVALID and b"temperature" in err()
    d.temperature = 0.02
    assert loss(4, 4, dim=200) == UNSUPPORTED and b"200" in err()
    assert bi_loss(4, 4) == INVALID and b"normalize_scores" in err()                            # no query rows to count
    d.normalize_scores, d.mode = 0, 2
    assert bi_loss(4, 6) == INVALID and b"multiple of n_queries" in err()                       # BiSigmoidLoss block walk
    d.mode = 3
    assert bi_loss(4, 8) == INVALID and b"square" in err()                                      # BiPairedEncoderLoss
    d.mode = 0
    assert bi_loss(4, 8) == INVALID and b"null device pointer" in err()                         # valid shape, no buffers
    d.struct_size = 8
    assert loss(4, 4) == INVALID and b"struct_size" in err()
    # dense dot products
    dd = DenseDotArgs(m=4, n=0, k=8)
    assert lib.cpb_dense_dot_launch(ctypes.byref(dd)) == INVALID and b"positive" in err()
    dd.n = 4
    assert lib.cpb_dense_dot_launch(ctypes.byref(dd)) == INVALID and b"null device pointer" in err()
    dd.d_a = dd.d_b = dd.d_out = 16
    dd.out_row_stride = 2
    assert lib.cpb_dense_dot_launch(ctypes.byref(dd)) == INVALID and b"out_row_stride" in err()
    assert lib.cpb_dense_dot_launch(None) == INVALID
    # forward
    a = MaxSimArgs(n_queries=1, nq_pad=32, n_docs=1, dim=128)
    assert lib.cpb_maxsim_launch(ctypes.byref(a)) == INVALID and b"null device pointer" in err()
    a.dim = 100
    assert lib.cpb_maxsim_launch(ctypes.byref(a)) == UNSUPPORTED and b"100" in err()
    a.dim, a.smooth_tau = 128, -1.0
    a.d_q = a.d_docs = a.d_doc_start = a.d_doc_len = a.d_scores = 16  # non-null, never dereferenced on the host
    a.doc_rows = 10
    assert lib.cpb_maxsim_launch(ctypes.byref(a)) == INVALID and b"smooth_tau" in err()
    a.smooth_tau, a.nq_real = 0.1, 0
    assert lib.cpb_maxsim_launch(ctypes.byref(a)) == INVALID and b"nq_real" in err()
    a.struct_size = 16
    assert lib.cpb_maxsim_launch(ctypes.byref(a)) == INVALID and b"struct_size" in err()
    assert lib.cpb_maxsim_launch(None) == INVALID
    # backward
    b = MaxSimBwdArgs(n_queries=4, nq_pad=32, n_docs=3, dim=128, doc_rows=10, max_doc_len=5)
    assert lib.cpb_maxsim_bwd_launch(ctypes.byref(b)) == INVALID and b"exactly one of" in err()
    b.d_argmax = 16
    assert lib.cpb_maxsim_bwd_launch(ctypes.byref(b)) == INVALID and b"null device pointer" in err()
    b.d_argmax, b.d_lse, b.dim = None, 16, 320
    b.d_grad_scores = b.d_q = b.d_docs = b.d_doc_start = b.d_doc_len = 16
    assert lib.cpb_maxsim_bwd_launch(ctypes.byref(b)) == INVALID and b"smooth_tau" in err()
    b.smooth_tau = 0.1
    assert lib.cpb_maxsim_bwd_launch(ctypes.byref(b)) == UNSUPPORTED and b"dim 128 only" in err()
    # head
    rc = lib.cpb_head_fwd(None, 0, 1536, None, None, 128, None, None, None, 0, None)
    assert rc == INVALID
    rc = lib.cpb_head_fwd(None, 10, 1536, None, None, 352, None, None, None, 0, None)
    assert rc == UNSUPPORTED and b"352" in err()                               # above ColQwen3's 320
    rc = lib.cpb_head_fwd(None, 10, 1536, None, None, 200, None, None, None, 0, None)
    assert rc == UNSUPPORTED and b"200" in err()                               # not a multiple of 32
    rc = lib.cpb_head_fwd(None, 10, 1000, None, None, 128, None, None, None, 0, None)
    assert rc == UNSUPPORTED and b"multiple of 64" in err()
    # helpers of the balanced / all-gather paths
    assert lib.cpb_wait_flags(None, 2, 1, None, None) == INVALID
    assert lib.cpb_maxsim_split_workspace_bytes(32, 32) > 0
    # tuning knobs
    assert lib.cpb_set_option(b"cluster", 3) == INVALID and lib.cpb_set_option(b"no_such_option", 1) == INVALID
    assert lib.cpb_set_option(b"cluster", 0) == 0 and lib.cpb_set_option(b"balanced", 1) == 0
    assert lib.cpb_set_option(b"pdl", 2) == INVALID and lib.cpb_set_option(b"pdl", 1) == 0
