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
    "cpb_colbert_loss_launch",
    "cpb_maxsim_launch",
    "cpb_maxsim_fwd",
    "cpb_maxsim_workspace_bytes",
    "cpb_maxsim_split_workspace_bytes",
    "cpb_maxsim_topk_workspace_bytes",
    "cpb_wait_flags",
    "cpb_maxsim_bwd_launch",
    "cpb_exchange_push",
    "cpb_signal_peers",
    "cpb_head_fwd",
    "cpb_dense_dot_launch",
)

CPB_ABI_VERSION = 3
CPB_FLAG_ROUND_BF16 = 1
CPB_FLAG_CONTIGUOUS = 2
CPB_FLAG_INDEPENDENT = 4
CPB_FLAG_GRAD_BF16 = 8
CPB_HEAD_CLAMP_NORM = 1
CPB_HEAD_SINGLE_ROUNDING = 2
CPB_TOPK_MAX = 16
CPB_LOSS_CE = 0
CPB_LOSS_PAIRWISE = 1
CPB_LOSS_SIGMOID = 2
CPB_LOSS_SYMMETRIC_CE = 3
CPB_DOT_A_F32 = 1
CPB_DOT_ACCUMULATE = 2
CPB_DOT_B_F32 = 4

c_vp, c_i, c_i64, c_u32, c_u64, c_f = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64,
                                       ctypes.c_float)


class LossDesc(ctypes.Structure):
    """``cpb_loss_desc`` (include/colpali_b200.h)."""

    _fields_ = [
        ("struct_size", c_u32), ("mode", c_i), ("normalize_scores", c_i), ("pos_aware_negative_filtering", c_i),
        ("offset", c_i), ("temperature", c_f), ("filter_threshold", c_f), ("filter_factor", c_f),
        ("d_neg_scores", c_vp), ("n_neg", c_i), ("in_batch_term_weight", c_f),
        ("d_loss", c_vp), ("d_grad_scores", c_vp), ("d_grad_neg_scores", c_vp), ("d_bounds", c_vp),
        ("neg_pos_offset_delta", c_i),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = ctypes.sizeof(LossDesc)


class MaxSimArgs(ctypes.Structure):
    """``cpb_maxsim_args`` (include/colpali_b200.h)."""

    _fields_ = [
        ("struct_size", c_u32), ("flags", c_u32), ("stream", c_vp),
        ("d_q", c_vp), ("n_queries", c_i), ("nq_pad", c_i), ("nq_real", c_i), ("dim", c_i),
        ("d_docs", c_vp), ("doc_rows", c_i64), ("d_doc_start", c_vp), ("d_doc_len", c_vp), ("d_doc_floor", c_vp),
        ("n_docs", c_i), ("uniform_len", c_i), ("max_doc_len", c_i),
        ("d_scores", c_vp), ("d_argmax", c_vp), ("d_lse", c_vp), ("d_workspace", c_vp),
        ("d_split_ws", c_vp), ("split_ws_bytes", c_i64), ("epoch", c_u32),
        ("smooth_tau", c_f),
        ("d_peer_bases", c_vp), ("mc_base", c_u64), ("n_peers", c_i), ("slab_word_offset", c_i64),
        ("flag_word_offset", c_i64),
        ("d_wait_flags", c_vp), ("n_wait", c_i), ("wait_value", c_u32),
        ("loss", ctypes.POINTER(LossDesc)), ("d_done_counter", c_vp),
        ("grid_out", c_i),
        ("d_topk_scores", c_vp), ("d_topk_idx", c_vp), ("d_topk_ws", c_vp), ("topk_k", c_i),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = ctypes.sizeof(MaxSimArgs)


class MaxSimBwdArgs(ctypes.Structure):
    """``cpb_maxsim_bwd_args`` (include/colpali_b200.h)."""

    _fields_ = [
        ("struct_size", c_u32), ("flags", c_u32), ("stream", c_vp),
        ("d_grad_scores", c_vp), ("d_grad_out", c_vp), ("d_argmax", c_vp), ("d_lse", c_vp), ("smooth_tau", c_f),
        ("d_q", c_vp), ("n_queries", c_i), ("nq_pad", c_i), ("nq_real", c_i), ("dim", c_i),
        ("d_docs", c_vp), ("doc_rows", c_i64), ("d_doc_start", c_vp), ("d_doc_len", c_vp), ("n_docs", c_i),
        ("max_doc_len", c_i),
        ("d_dq", c_vp), ("d_dd", c_vp), ("d_dd_doc_base", c_vp),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = ctypes.sizeof(MaxSimBwdArgs)


class ExchangePushArgs(ctypes.Structure):
    """``cpb_exchange_push_args`` (include/colpali_b200.h)."""

    _fields_ = [
        ("struct_size", c_u32), ("pad_first", c_u32), ("stream", c_vp),
        ("d_src", c_vp), ("n_docs", c_i), ("len", c_i), ("slot_len", c_i), ("dim", c_i),
        ("d_peer_bases", c_vp), ("mc_base", c_u64), ("n_peers", c_i),
        ("bank_word_offset", c_i64), ("flag_word_offset", c_i64), ("grid_out", c_i),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = ctypes.sizeof(ExchangePushArgs)


class DenseDotArgs(ctypes.Structure):
    """``cpb_dense_dot_args`` (include/colpali_b200.h)."""

    _fields_ = [
        ("struct_size", c_u32), ("flags", c_u32), ("stream", c_vp),
        ("d_a", c_vp), ("a_row_stride", c_i64), ("a_k_stride", c_i64),
        ("d_b", c_vp), ("b_row_stride", c_i64), ("b_k_stride", c_i64), ("d_b_rows", c_vp),
        ("m", c_i), ("n", c_i), ("k", c_i),
        ("d_out", c_vp), ("out_row_stride", c_i64), ("d_alpha", c_vp),
    ]

    def __init__(self, **kw):
        super().__init__(**kw)
        self.struct_size = ctypes.sizeof(DenseDotArgs)


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
    ci = ctypes.c_int

    lib.cpb_abi_version.restype = ci
    lib.cpb_abi_version.argtypes = []
    if lib.cpb_abi_version() != CPB_ABI_VERSION:
        raise ColpaliB200Error(f"{LIB_PATH} has ABI version {lib.cpb_abi_version()}, this package needs {CPB_ABI_VERSION}: "
                               "rebuild with `python -m colpali_b200.build --force`")
    lib.cpb_last_error.restype = ctypes.c_char_p
    lib.cpb_last_error.argtypes = []
    lib.cpb_device_info.restype = ci
    lib.cpb_device_info.argtypes = [ci, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    lib.cpb_set_option.restype = ci
    lib.cpb_set_option.argtypes = [ctypes.c_char_p, ci]
    lib.cpb_maxsim_workspace_bytes.restype = c_i64
    lib.cpb_maxsim_workspace_bytes.argtypes = [ci, ci, ci]
    lib.cpb_maxsim_split_workspace_bytes.restype = c_i64
    lib.cpb_maxsim_split_workspace_bytes.argtypes = [ci, ci]
    lib.cpb_maxsim_topk_workspace_bytes.restype = c_i64
    lib.cpb_maxsim_topk_workspace_bytes.argtypes = []
    lib.cpb_maxsim_launch.restype = ci
    lib.cpb_maxsim_launch.argtypes = [ctypes.POINTER(MaxSimArgs)]
    lib.cpb_maxsim_fwd.restype = ci
    lib.cpb_maxsim_fwd.argtypes = [
        c_vp, ci, ci,  # d_q, n_queries, nq_pad
        c_vp, c_i64,  # d_docs, doc_rows
        c_vp, c_vp, c_vp, ci,  # d_doc_start, d_doc_len, d_doc_floor, n_docs
        c_vp, c_vp, c_vp,  # d_scores, d_argmax, d_workspace
        c_u32, c_vp,  # flags, stream
    ]
    lib.cpb_colbert_loss_launch.restype = ci
    lib.cpb_colbert_loss_launch.argtypes = [ctypes.POINTER(LossDesc), c_vp, c_vp, ci, ci, ci, ci, c_vp]
    lib.cpb_wait_flags.restype = ci
    lib.cpb_wait_flags.argtypes = [c_vp, ci, c_u32, c_vp, c_vp]
    lib.cpb_maxsim_bwd_launch.restype = ci
    lib.cpb_maxsim_bwd_launch.argtypes = [ctypes.POINTER(MaxSimBwdArgs)]
    lib.cpb_exchange_push.restype = ci
    lib.cpb_exchange_push.argtypes = [ctypes.POINTER(ExchangePushArgs)]
    lib.cpb_signal_peers.restype = ci
    lib.cpb_signal_peers.argtypes = [c_vp, c_u64, ci, c_i64, c_vp]
    lib.cpb_dense_dot_launch.restype = ci
    lib.cpb_dense_dot_launch.argtypes = [ctypes.POINTER(DenseDotArgs)]
    lib.cpb_head_fwd.restype = ci
    lib.cpb_head_fwd.argtypes = [
        c_vp, c_i64, ci,  # d_hidden, n_tokens, hidden
        c_vp, c_vp, ci,  # d_weight, d_bias, dim
        c_vp, c_vp,  # d_attention_mask, d_extra_mask
        c_vp, c_u32, c_vp,  # d_out, flags, stream
    ]
    _lib = lib
    # experiment switch: COLPALI_B200_OPTS="pdl=0,boundary_mode=0" applies cpb_set_option pairs at load time
    for pair in filter(None, os.environ.get("COLPALI_B200_OPTS", "").split(",")):
        name, _, value = pair.partition("=")
        if lib.cpb_set_option(name.strip().encode(), int(value)) != 0:
            raise ColpaliB200Error(f"COLPALI_B200_OPTS: {lib.cpb_last_error().decode()}")
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cpb_last_error().decode("utf-8", "replace")
        raise ColpaliB200Error(f"{what} failed (code {rc}): {msg}")


def set_option(name: str, value: int) -> None:
    """Tuning knob passthrough (cpb_set_option), e.g. 'cluster', 'qtiles_per_cta', 'balanced', 'pdl'."""
    check(load().cpb_set_option(name.encode(), int(value)), f"cpb_set_option({name})")


def gpu_launches() -> int:
    """Number of kernels this process launched through the C ABI (bench.py reports it)."""
    return _launch_count


_launch_count = 0


def count_launches(n: int) -> None:
    global _launch_count
    _launch_count += n
