"""ctypes binding of libtavsr_sm100.so (C ABI declared in include/tavsr.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, the product
path raises.  `python -m tailored_avsr_b200.build` (or `__graft_entry__.build()`) builds it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong,
                    c_size_t, c_void_p)

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libtavsr_sm100.so")

ACT_NONE, ACT_SWISH, ACT_GELU, ACT_RELU = 0, 1, 2, 3
DT_TF32, DT_BF16 = 0, 1
DT_OUT_BF16, DT_LNA_BF16, DT_LNB_BF16 = 0x100, 0x200, 0x400


class TavsrError(RuntimeError):
    """Raised when a libtavsr_sm100 entry point reports an error."""


class RowLNArgs(Structure):
    """Mirror of `tavsr_rowln_args` (include/tavsr.h)."""

    _fields_ = [
        ("struct_size", c_int), ("M", c_int), ("K", c_int), ("dtype", c_int),
        ("x", c_void_p), ("ldx", c_longlong),
        ("x2", c_void_p), ("ldx2", c_longlong),
        ("w", c_void_p), ("ldw", c_longlong),
        ("bias", c_void_p),
        ("residual", c_void_p), ("ldr", c_longlong), ("alpha", c_float),
        ("rowscale1", c_void_p), ("rowscale2", c_void_p), ("rows_per_seg", c_int),
        ("ln0_g", c_void_p), ("ln0_b", c_void_p), ("eps0", c_float),
        ("out_main", c_void_p), ("ld_main", c_longlong), ("round_main", c_int),
        ("lnA_g", c_void_p), ("lnA_b", c_void_p), ("out_lnA", c_void_p), ("ld_lnA", c_longlong),
        ("round_lnA", c_int),
        ("lnB_g", c_void_p), ("lnB_b", c_void_p), ("out_lnB", c_void_p), ("ld_lnB", c_longlong),
        ("round_lnB", c_int),
        ("eps", c_float),
        ("dot1", c_void_p), ("dot2", c_void_p), ("dots_out", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_longlong),
        ("k1", c_int), ("segbias1", c_void_p), ("segbias2", c_void_p),
    ]


class FfnArgs(Structure):
    """Mirror of `tavsr_ffn_args` (include/tavsr.h)."""

    _fields_ = [
        ("struct_size", c_int), ("hidden", c_int), ("act", c_int), ("reserved", c_int),
        ("xn", c_void_p), ("ldxn", c_longlong),
        ("w1", c_void_p), ("ldw1", c_longlong),
        ("b1", c_void_p),
        ("w2", c_void_p), ("ldw2", c_longlong),
        ("ep", RowLNArgs),
    ]


# name -> (restype, argtypes); every symbol include/tavsr.h declares
SIGNATURES = {
    "tavsr_version": (c_int, []),
    "tavsr_last_error": (c_char_p, []),
    "tavsr_debug_set": (c_int, [c_int, c_int]),
    "tavsr_debug_set_ptr": (c_int, [c_void_p]),
    "tavsr_launch_count": (c_longlong, []),
    "tavsr_gemm_bias_act": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                    c_longlong, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    "tavsr_gemm_group2": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                  c_longlong, c_int, c_void_p, c_longlong, c_void_p, c_longlong,
                                  c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_int,
                                  c_void_p]),
    "tavsr_gemm_rowln": (c_int, [POINTER(RowLNArgs), c_void_p]),
    "tavsr_rowln_workspace_bytes": (c_size_t, [c_int]),
    "tavsr_ffn_fused": (c_int, [POINTER(FfnArgs), c_void_p]),
    "tavsr_layernorm": (c_int, [c_void_p, c_longlong, c_int, c_int, c_float, c_void_p, c_void_p,
                                c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p,
                                c_longlong, c_int, c_float, c_int, c_void_p]),
    "tavsr_relpos_attn_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p]),
    "tavsr_gemm_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tavsr_gemm_wgrad": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                 c_int, c_int, c_int, c_void_p, c_longlong, c_void_p]),
    "tavsr_relpos_attn_fwd_dropout": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int,
                                              c_int, c_void_p, c_void_p, c_longlong, c_float,
                                              c_void_p]),
    "tavsr_relpos_attn_bwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                      c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_longlong, c_float, c_int, c_int, c_int, c_void_p]),
    "tavsr_conv2d_sub_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tavsr_conv2d_sub_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                     c_void_p, c_void_p, c_longlong, c_void_p]),
    "tavsr_act_fwd_t": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int,
                                c_int, c_int, c_void_p]),
    "tavsr_act_bwd_t": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
                                c_void_p, c_longlong, c_int, c_int, c_int, c_void_p]),
    "tavsr_act_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int,
                              c_int, c_int, c_void_p]),
    "tavsr_merge_learned_ave_weights2": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p,
                                                 c_void_p, c_float, c_float, c_float, c_float,
                                                 c_float, c_void_p, c_void_p, c_int, c_int,
                                                 c_void_p]),
    "tavsr_scale_add_rows": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                     c_void_p, c_int, c_void_p, c_longlong, c_int, c_int, c_int,
                                     c_void_p]),
    "tavsr_csgu_fwd": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_int,
                               c_float, c_int, c_int, c_void_p]),
    "tavsr_row_dots": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               c_void_p]),
    "tavsr_merge_learned_ave_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float,
                                                c_float, c_float, c_float, c_void_p, c_void_p,
                                                c_int, c_int, c_void_p]),
    "tavsr_split_tf32": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int,
                                 c_void_p]),
    "tavsr_conv2d_sub_im2col": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                        c_void_p, c_int, c_void_p]),
    "tavsr_ctc_head": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "tavsr_vocab_residual": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_longlong, c_void_p, c_void_p, c_float, c_void_p, c_longlong,
                                     c_int, c_int, c_int, c_int, c_void_p]),
    "tavsr_ctc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tavsr_ctc_loss": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_void_p]),
    "tavsr_ctc_head_bwd_workspace_bytes": (c_size_t, [c_int]),
    "tavsr_ctc_head_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_longlong, c_void_p, c_void_p,
                                   c_longlong, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int,
                                   c_int, c_void_p]),
    "tavsr_transpose_2d": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "tavsr_col_sums_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tavsr_col_sums": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                               c_longlong, c_int, c_int, c_void_p]),
    "tavsr_act_bwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong, c_int,
                              c_int, c_int, c_void_p]),
    "tavsr_layernorm_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tavsr_layernorm_bwd": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_longlong, c_void_p,
                                    c_longlong, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                    c_longlong, c_int, c_int, c_float, c_void_p]),
    "tavsr_csgu_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tavsr_csgu_conv_bwd": (c_int, [c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                    c_longlong, c_void_p, c_void_p, c_void_p, c_longlong, c_int,
                                    c_int, c_int, c_int, c_void_p]),
    "tavsr_merge_learned_ave_bwd_workspace_bytes": (c_size_t, [c_int]),
    "tavsr_merge_learned_ave_bwd": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p,
                                            c_longlong, c_void_p, c_void_p,            # lens, lens2
                                            c_void_p, c_void_p, c_void_p, c_void_p,    # a1 b1 a2 b2
                                            c_void_p, c_void_p, c_longlong, c_void_p, c_longlong,
                                            c_void_p, c_void_p, c_longlong, c_int, c_int, c_int,
                                            c_void_p]),
    "tavsr_merge_scores": (c_int, [c_void_p, c_longlong, c_int, c_void_p, c_longlong, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                   c_float, c_float, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_void_p]),
    "tavsr_merge_learned_ave_weights_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                                    c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "tavsr_softmax_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "tavsr_ctc_greedy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_void_p]),
    "tavsr_ctc_prefix_score": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and declare every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TavsrError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no "
            "fallback path.  Run `python -m tailored_avsr_b200.build`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    # debug knobs, e.g. TAVSR_DEBUG="3=1" forces the 1-CTA GEMM (see gemm_sm100.cu)
    for kv in filter(None, os.environ.get("TAVSR_DEBUG", "").split(",")):
        k, v = kv.split("=")
        lib.tavsr_debug_set(int(k), int(v))
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tavsr_last_error()
        raise TavsrError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
