"""ctypes binding of libcolpali_b200.so (the C ABI declared in include/colpali_b200.h).

There is deliberately no CPU or PyTorch fallback: if the shared library has not been built, or a
call fails, the product path raises.  Build with ``python -m colpali_b200.build`` (or
``__graft_entry__.build()``).
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COLPALI_B200_LIB") or os.path.join(_HERE, "libcolpali_b200.so")  # env override: kernel-variant experiments

# every symbol include/colpali_b200.h declares; tests check the built library exports them all
EXPORTED_SYMBOLS = (
    "cpb_abi_version",
    "cpb_last_error",
    "cpb_device_info",
    "cpb_set_option",
    "cpb_maxsim_fwd",
    "cpb_maxsim_workspace_bytes",
    "cpb_maxsim_fwd_balanced",
    "cpb_maxsim_fwd_dim",
    "cpb_maxsim_split_workspace_bytes",
    "cpb_maxsim_fwd_allgather",
    "cpb_wait_flags",
    "cpb_colbert_loss_fwd",
    "cpb_colbert_neg_loss_fwd",
    "cpb_maxsim_bwd",
    "cpb_maxsim_bwd_dim",
    "cpb_colbert_loss_fwd_dim",
    "cpb_colbert_neg_loss_fwd_dim",
    "cpb_head_fwd",
)

CPB_FLAG_ROUND_BF16 = 1
CPB_FLAG_CONTIGUOUS = 2
CPB_HEAD_CLAMP_NORM = 1
CPB_HEAD_SINGLE_ROUNDING = 2
CPB_LOSS_CE = 0
CPB_LOSS_PAIRWISE = 1
CPB_LOSS_SIGMOID = 2

_lib: Optional[ctypes.CDLL] = None


class ColpaliB200Error(RuntimeError):
    """A libcolpali_b200 call returned a non-zero status."""


def load() -> ctypes.CDLL:
    """Load (once) and return the shared library.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ColpaliB200Error(
            f"{LIB_PATH} not found: the sm_100a extension is not built. "
            "Run `python -m colpali_b200.build` (needs nvcc). There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    c_vp, c_i, c_i64, c_u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32

    lib.cpb_abi_version.restype = c_i
    lib.cpb_abi_version.argtypes = []
    lib.cpb_last_error.restype = ctypes.c_char_p
    lib.cpb_last_error.argtypes = []
    lib.cpb_device_info.restype = c_i
    lib.cpb_device_info.argtypes = [c_i, ctypes.POINTER(c_i), ctypes.POINTER(c_i), ctypes.POINTER(c_i)]
    lib.cpb_set_option.restype = c_i
    lib.cpb_set_option.argtypes = [ctypes.c_char_p, c_i]
    lib.cpb_maxsim_workspace_bytes.restype = c_i64
    lib.cpb_maxsim_workspace_bytes.argtypes = [c_i, c_i, c_i]
    lib.cpb_maxsim_fwd.restype = c_i
    lib.cpb_maxsim_fwd.argtypes = [
        c_vp, c_i, c_i,  # d_q, n_queries, nq_pad
        c_vp, c_i64,  # d_docs, doc_rows
        c_vp, c_vp, c_vp, c_i,  # d_doc_start, d_doc_len, d_doc_floor, n_docs
        c_vp, c_vp, c_vp,  # d_scores, d_argmax, d_workspace
        c_u32, c_vp,  # flags, stream
    ]
    lib.cpb_maxsim_fwd_dim.restype = c_i
    lib.cpb_maxsim_fwd_dim.argtypes = lib.cpb_maxsim_fwd.argtypes[:-1] + [c_i, c_vp]
    lib.cpb_maxsim_split_workspace_bytes.restype = c_i64
    lib.cpb_maxsim_split_workspace_bytes.argtypes = [c_i, c_i]
    lib.cpb_maxsim_fwd_balanced.restype = c_i
    lib.cpb_maxsim_fwd_balanced.argtypes = lib.cpb_maxsim_fwd.argtypes[:-1] + [c_i, c_i, c_vp, c_i64, c_u32, c_vp]
    lib.cpb_maxsim_fwd_allgather.restype = c_i
    lib.cpb_maxsim_fwd_allgather.argtypes = [
        c_vp, c_i, c_i, c_vp, c_i64, c_vp, c_vp, c_vp, c_i,  # q, n_queries, nq_pad, docs, rows, start, len, floor, n_docs
        c_vp, c_i, c_i, c_u32,  # d_peer_slabs, n_peers, my_rank, flags
        c_i, c_i, c_vp, c_i64, c_u32,  # uniform_len, max_doc_len, d_split_ws, split_ws_bytes, epoch
        c_vp, c_i64, c_u32, c_vp,  # d_done_counter, flag_word_offset, signal_value, stream
    ]
    lib.cpb_wait_flags.restype = c_i
    lib.cpb_wait_flags.argtypes = [c_vp, c_i, c_u32, c_vp]
    c_f = ctypes.c_float
    lib.cpb_colbert_loss_fwd.restype = c_i
    lib.cpb_colbert_loss_fwd.argtypes = [
        c_vp, c_vp, c_i, c_i, c_i, c_i,  # d_scores, d_q, n_queries, nq_pad, n_docs, mode
        c_f, c_i, c_i, c_f, c_f, c_i,  # temperature, normalize, filter, threshold, factor, offset
        c_vp, c_vp, c_vp, c_vp,  # d_loss, d_grad_scores, d_bounds, stream
    ]
    lib.cpb_colbert_neg_loss_fwd.restype = c_i
    lib.cpb_colbert_neg_loss_fwd.argtypes = [
        c_vp, c_vp, c_vp, c_i, c_i, c_i, c_i, c_i,  # d_scores, d_neg_scores, d_q, n_queries, nq_pad, n_docs, n_neg, inner_mode
        c_f, c_i, c_i, c_f, c_f, c_f, c_i,  # temperature, normalize, filter, threshold, factor, in_batch_weight, offset
        c_vp, c_vp, c_vp, c_vp,  # d_loss, d_grad_scores, d_grad_neg_scores, stream
    ]
    for name in ("cpb_colbert_loss_fwd", "cpb_colbert_neg_loss_fwd"):  # DRAFT: ..., dim, stream
        fn = getattr(lib, name + "_dim")
        fn.restype = c_i
        fn.argtypes = getattr(lib, name).argtypes[:-1] + [c_i, c_vp]
    lib.cpb_maxsim_bwd.restype = c_i
    lib.cpb_maxsim_bwd.argtypes = [
        c_vp, c_vp, c_vp,  # d_grad_scores, d_grad_out, d_argmax
        c_vp, c_i, c_i,  # d_q, n_queries, nq_pad
        c_vp, c_i64, c_vp, c_i,  # d_docs, doc_rows, d_doc_start, n_docs
        c_vp, c_vp, c_vp,  # d_dq, d_dd, stream
    ]
    lib.cpb_maxsim_bwd_dim.restype = c_i
    lib.cpb_maxsim_bwd_dim.argtypes = lib.cpb_maxsim_bwd.argtypes[:-1] + [c_i, c_vp]  # ..., dim, stream
    lib.cpb_head_fwd.restype = c_i
    lib.cpb_head_fwd.argtypes = [
        c_vp, c_i64, c_i,  # d_hidden, n_tokens, hidden
        c_vp, c_vp, c_i,  # d_weight, d_bias, dim
        c_vp, c_vp,  # d_attention_mask, d_extra_mask
        c_vp, c_u32, c_vp,  # d_out, flags, stream
    ]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cpb_last_error().decode("utf-8", "replace")
        raise ColpaliB200Error(f"{what} failed (code {rc}): {msg}")


def set_option(name: str, value: int) -> None:
    """Tuning knob passthrough (cpb_set_option): 'cluster', 'qtiles_per_cta', 'debug_flags'."""
    check(load().cpb_set_option(name.encode(), int(value)), f"cpb_set_option({name})")


def gpu_launches() -> int:
    """Number of kernels this process launched through the C ABI (bench.py reports it)."""
    return _launch_count


_launch_count = 0


def count_launches(n: int) -> None:
    global _launch_count
    _launch_count += n
