"""Tensor-level wrappers over the C ABI (include/tavsr.h).

PyTorch is used only for device memory and streams: every function here takes CUDA tensors, passes
raw device pointers + the current stream to libtavsr_sm100.so and returns the output tensors.
There is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SWISH, DT_BF16, DT_LNA_BF16, DT_LNB_BF16,
                   DT_OUT_BF16, DT_TF32, FfnArgs, RowLNArgs, check)

__all__ = [
    "ACT_NONE", "ACT_SWISH", "ACT_GELU", "ACT_RELU",
    "gemm_bias_act", "gemm_group2", "conv2d_sub_im2col", "merge_weights2", "scale_add_rows", "cast_bf16", "split_tf32", "gemm_rowln", "ffn_fused", "layernorm", "relpos_attn", "csgu", "merge_weights", "merge_scores", "merge_weights_dev",
    "ctc_head", "ctc_head_bwd", "vocab_residual", "row_dots", "ctc_loss", "ctc_greedy", "ctc_prefix_score", "launch_count",
]


_PROFILE = None  # list collecting (op name, arg shapes, start event, end event) when profiling


def set_profiler(records) -> None:
    """bench.py hook: pass a list to time every op with CUDA events on the launching stream
    (None switches it off)."""
    global _PROFILE
    _PROFILE = records


def _profiled(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        # launch on the device (and its current stream) that owns the operands, not on whatever
        # device happens to be current
        t0 = next((a for a in args if torch.is_tensor(a) and a.is_cuda), None)
        if t0 is not None and t0.device.index != torch.cuda.current_device():
            with torch.cuda.device(t0.device):
                return wrapper(*args, **kwargs)
        if _PROFILE is None:
            return fn(*args, **kwargs)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args, **kwargs)
        e1.record()
        shapes = tuple(tuple(a.shape) for a in args if torch.is_tensor(a))
        extra = {k: (tuple(v.shape) if torch.is_tensor(v) else v) for k, v in kwargs.items()
                 if k in ("act", "x2", "ln0", "lnA", "lnB", "residual") and v is not None}
        _PROFILE.append((fn.__name__, shapes, extra, e0, e1))
        return out

    return wrapper


try:   # the raw handle without building a torch.cuda.Stream object per launch (~10 us each)
    _raw_stream = torch._C._cuda_getCurrentRawStream
    _cur_device = torch._C._cuda_getDevice
except AttributeError:  # pragma: no cover - older / CPU-only torch builds
    _raw_stream = _cur_device = None


def _stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_ACT_DTYPES = (torch.float32, torch.bfloat16)


def _chk2d(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    """dtype=None accepts either activation storage type (fp32 / bf16)."""
    if not t.is_cuda:
        raise _lib.TavsrError(f"{name} must be a CUDA tensor (no CPU fallback exists)")
    ok = t.dtype in _ACT_DTYPES if dtype is None else t.dtype == dtype
    if not ok or t.dim() != 2 or t.stride(1) != 1:
        raise _lib.TavsrError(f"{name} must be 2-D {dtype or 'fp32/bf16'} with unit inner stride, got "
                              f"{tuple(t.shape)} {t.dtype} strides {t.stride()}")


def _is_bf16(t: Optional[torch.Tensor]) -> bool:
    return t is not None and t.dtype == torch.bfloat16


def _chk_k(w: torch.Tensor, K: int, who: str) -> None:
    """The kernels take the reduction length from the activation operand: a weight of another
    width would be read out of bounds, so refuse it here."""
    if w.shape[1] != K:
        raise _lib.TavsrError(f"{who}: weight reduction width {w.shape[1]} != activation width {K}")


_SPLITK_WS = {}  # (device index, stream handle) -> zero-initialised split-K scratch


def _splitk_workspace(device, M: int) -> torch.Tensor:
    """Per-stream scratch for the split-K row-complete GEMM (flag words must start at zero; the
    kernel re-zeroes them, so one allocation serves every later launch on that stream)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    need = int(_lib.load().tavsr_rowln_workspace_bytes(M))
    ws = _SPLITK_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _SPLITK_WS[key] = ws
    return ws


def launch_count() -> int:
    return int(_lib.load().tavsr_launch_count())


@_profiled
def gemm_bias_act(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act: int = ACT_NONE,
                  round_out: bool = False, out: Optional[torch.Tensor] = None,
                  out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """out = act(x @ w.T + bias) on the tcgen05 GEMM (tavsr_gemm_bias_act).  fp32 operands run as
    TF32; bf16 operands (both) run kind::f16 MMAs with fp32 accumulate and an fp32 or bf16 output
    (`out_dtype`, default fp32, or the dtype of `out`)."""
    bf16 = x.dtype == torch.bfloat16
    _chk2d(x, "x", x.dtype if bf16 else torch.float32)
    _chk2d(w, "w", torch.bfloat16 if bf16 else torch.float32)
    M, K = x.shape
    N = w.shape[0]
    _chk_k(w, K, "gemm_bias_act")
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype or torch.float32)
    _chk2d(out, "out", None)
    if out.dtype == torch.bfloat16 and not bf16:
        raise _lib.TavsrError("gemm_bias_act: a bf16 output needs bf16 operands")
    dt = (DT_BF16 if bf16 else DT_TF32) | (DT_OUT_BF16 if out.dtype == torch.bfloat16 else 0)
    lib = _lib.load()
    check(lib.tavsr_gemm_bias_act(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), _p(bias),
                                  out.data_ptr(), out.stride(0), M, N, K, act, int(round_out),
                                  dt, _stream()), "tavsr_gemm_bias_act")
    return out


@_profiled
def gemm_group2(x1: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], x2: torch.Tensor,
                w2: torch.Tensor, b2: Optional[torch.Tensor],
                out_dtype: Optional[torch.dtype] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(x1 @ w1.T + b1, gelu(x2 @ w2.T + b2)) in one launch (tavsr_gemm_group2): the fused QKV
    projection and channel_proj1 + GELU of a two-branch block."""
    bf16 = x1.dtype == torch.bfloat16
    for t, n in ((x1, "x1"), (x2, "x2"), (w1, "w1"), (w2, "w2")):
        _chk2d(t, n, x1.dtype)
    M, K = x1.shape
    if x2.shape != (M, K):
        raise _lib.TavsrError("gemm_group2: both problems must share M and K")
    _chk_k(w1, K, "gemm_group2 (w1)")
    _chk_k(w2, K, "gemm_group2 (w2)")
    odt = out_dtype or torch.float32
    y1 = torch.empty((M, w1.shape[0]), device=x1.device, dtype=odt)
    y2 = torch.empty((M, w2.shape[0]), device=x1.device, dtype=odt)
    dt = (DT_BF16 if bf16 else DT_TF32) | (DT_OUT_BF16 if odt == torch.bfloat16 else 0)
    check(_lib.load().tavsr_gemm_group2(
        x1.data_ptr(), x1.stride(0), w1.data_ptr(), w1.stride(0), _p(b1), y1.data_ptr(), y1.stride(0),
        w1.shape[0], x2.data_ptr(), x2.stride(0), w2.data_ptr(), w2.stride(0), _p(b2), y2.data_ptr(),
        y2.stride(0), w2.shape[0], M, K, dt, _stream()), "tavsr_gemm_group2")
    return y1, y2


def _fill_rowln(a: RowLNArgs, M: int, bias, residual, alpha, ln0, eps0, out_main, round_main, lnA,
                out_lnA, round_lnA, lnB, out_lnB, round_lnB, eps, dots, dots_out) -> None:
    a.struct_size = ctypes.sizeof(RowLNArgs)
    a.M = M
    a.dtype = ((DT_OUT_BF16 if _is_bf16(out_main) else 0) | (DT_LNA_BF16 if _is_bf16(out_lnA) else 0)
               | (DT_LNB_BF16 if _is_bf16(out_lnB) else 0))   # the caller ORs in the operand type
    a.bias = _p(bias)
    if residual is not None:
        _chk2d(residual, "residual")
        a.residual, a.ldr = residual.data_ptr(), residual.stride(0)
    a.alpha = alpha
    if ln0 is not None:
        a.ln0_g, a.ln0_b = ln0[0].data_ptr(), ln0[1].data_ptr()
    a.eps0 = eps0
    if out_main is not None:
        _chk2d(out_main, "out_main", None)
        a.out_main, a.ld_main = out_main.data_ptr(), out_main.stride(0)
    a.round_main = int(round_main)
    if lnA is not None:
        _chk2d(out_lnA, "out_lnA", None)
        a.lnA_g, a.lnA_b = lnA[0].data_ptr(), lnA[1].data_ptr()
        a.out_lnA, a.ld_lnA = out_lnA.data_ptr(), out_lnA.stride(0)
    a.round_lnA = int(round_lnA)
    if lnB is not None:
        _chk2d(out_lnB, "out_lnB", None)
        a.lnB_g, a.lnB_b = lnB[0].data_ptr(), lnB[1].data_ptr()
        a.out_lnB, a.ld_lnB = out_lnB.data_ptr(), out_lnB.stride(0)
    a.round_lnB = int(round_lnB)
    a.eps = eps
    if dots is not None:
        a.dot1, a.dot2 = dots[0].data_ptr(), dots[1].data_ptr()
        a.dots_out = dots_out.data_ptr()


@_profiled
def gemm_rowln(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
               x2: Optional[torch.Tensor] = None,
               rowscale: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, rows_per_seg: int = 0,
               residual: Optional[torch.Tensor] = None, alpha: float = 1.0,
               ln0: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eps0: float = 1e-12,
               out_main: Optional[torch.Tensor] = None, round_main: bool = False,
               lnA: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               out_lnA: Optional[torch.Tensor] = None, round_lnA: bool = False,
               lnB: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               out_lnB: Optional[torch.Tensor] = None, round_lnB: bool = False,
               eps: float = 1e-12,
               dots: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
               dots_out: Optional[torch.Tensor] = None, k1: int = 0,
               segbias: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> None:
    """Row-complete N=256 GEMM with fused residual / LayerNorm / row-dot epilogue (tavsr_gemm_rowln).
    k1 > 0 selects the sequential dual mode: x (M,k1), x2 (M,K-k1), w (256,K) = [W1 | W2]."""
    _chk2d(x, "x", None)
    _chk2d(w, "w", x.dtype)
    a = RowLNArgs()
    _fill_rowln(a, x.shape[0], bias, residual, alpha, ln0, eps0, out_main, round_main, lnA, out_lnA,
                round_lnA, lnB, out_lnB, round_lnB, eps, dots, dots_out)
    a.dtype |= DT_BF16 if _is_bf16(x) else DT_TF32
    a.K = x.shape[1]
    a.x, a.ldx = x.data_ptr(), x.stride(0)
    if x2 is not None:
        _chk2d(x2, "x2", x.dtype)
        a.x2, a.ldx2 = x2.data_ptr(), x2.stride(0)
        a.rowscale1, a.rowscale2 = rowscale[0].data_ptr(), rowscale[1].data_ptr()
        a.rows_per_seg = rows_per_seg
        if k1 > 0:
            a.K = x.shape[1] + x2.shape[1]
            a.k1 = k1
            if segbias is not None:
                a.segbias1, a.segbias2 = segbias[0].data_ptr(), segbias[1].data_ptr()
    a.w, a.ldw = w.data_ptr(), w.stride(0)
    if w.shape[0] != 256:
        raise _lib.TavsrError(f"gemm_rowln: the row-complete epilogue is built for N = 256, got {w.shape[0]}")
    if x2 is not None and k1 == 0:
        _chk_k(w, x.shape[1], "gemm_rowln")
        if x2.shape[1] != x.shape[1]:
            raise _lib.TavsrError("gemm_rowln: dual operands must have equal widths")
    else:
        _chk_k(w, a.K, "gemm_rowln")
    if x2 is None and a.K >= 1024:
        ws = _splitk_workspace(x.device, a.M)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    check(_lib.load().tavsr_gemm_rowln(ctypes.byref(a), _stream()), "tavsr_gemm_rowln")


@_profiled
def ffn_fused(xn: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
              b2: Optional[torch.Tensor], act: int, *,
              residual: Optional[torch.Tensor] = None, alpha: float = 1.0,
              ln0: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eps0: float = 1e-12,
              out_main: Optional[torch.Tensor] = None, round_main: bool = False,
              lnA: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              out_lnA: Optional[torch.Tensor] = None, round_lnA: bool = False,
              lnB: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              out_lnB: Optional[torch.Tensor] = None, round_lnB: bool = False,
              eps: float = 1e-12) -> None:
    """residual + alpha * (act(xn @ w1.T + b1) @ w2.T + b2) with the row-complete LayerNorm epilogue;
    the hidden activation stays on chip (tavsr_ffn_fused)."""
    _chk2d(xn, "xn", None)
    _chk2d(w1, "w1", xn.dtype)
    _chk2d(w2, "w2", xn.dtype)
    _chk_k(w1, xn.shape[1], "ffn_fused (w1)")
    _chk_k(w2, w1.shape[0], "ffn_fused (w2)")
    if xn.shape[1] != 256 or w2.shape[0] != 256:
        raise _lib.TavsrError("ffn_fused: built for model width 256")
    a = FfnArgs()
    a.struct_size = ctypes.sizeof(FfnArgs)
    a.hidden = w1.shape[0]
    a.act = act
    a.xn, a.ldxn = xn.data_ptr(), xn.stride(0)
    a.w1, a.ldw1 = w1.data_ptr(), w1.stride(0)
    a.b1 = _p(b1)
    a.w2, a.ldw2 = w2.data_ptr(), w2.stride(0)
    _fill_rowln(a.ep, xn.shape[0], b2, residual, alpha, ln0, eps0, out_main, round_main, lnA, out_lnA,
                round_lnA, lnB, out_lnB, round_lnB, eps, None, None)
    a.ep.dtype |= DT_BF16 if _is_bf16(xn) else DT_TF32
    check(_lib.load().tavsr_ffn_fused(ctypes.byref(a), _stream()), "tavsr_ffn_fused")


@_profiled
def layernorm(x: torch.Tensor, gA: torch.Tensor, bA: torch.Tensor, eps: float = 1e-12,
              round_out: bool = False, scale: float = 1.0, out: Optional[torch.Tensor] = None,
              gB: Optional[torch.Tensor] = None, bB: Optional[torch.Tensor] = None,
              outB: Optional[torch.Tensor] = None, roundB: bool = False,
              out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """LayerNorm of fp32 rows into an fp32 or bf16 output (`out_dtype` / the dtype of `out`)."""
    _chk2d(x, "x")
    M, D = x.shape
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=out_dtype or torch.float32)
    dt = (DT_LNA_BF16 if _is_bf16(out) else 0) | (DT_LNB_BF16 if _is_bf16(outB) else 0)
    check(_lib.load().tavsr_layernorm(x.data_ptr(), x.stride(0), M, D, eps, gA.data_ptr(),
                                      bA.data_ptr(), out.data_ptr(), out.stride(0), int(round_out),
                                      _p(gB), _p(bB), _p(outB),
                                      outB.stride(0) if outB is not None else 0, int(roundB),
                                      scale, dt, _stream()), "tavsr_layernorm")
    return out


@_profiled
def relpos_attn(qkv: torch.Tensor, pos: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                lens: Optional[torch.Tensor], B: int, T: int, H: int, round_out: bool = True,
                out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None,
                drop: Optional[tuple] = None):
    """ctx = rel-pos MHSA(qkv) for (B*T, 3*H*64) fused projections (tavsr_relpos_attn_fwd); fp32
    qkv / pos -> fp32 ctx on TF32 MMAs, bf16 qkv / pos -> bf16 ctx on kind::f16 MMAs.  `lse`
    (B, H, T) fp32, optional: receives the per-row log-sum-exp the backward recomputes P from.
    `drop` = (keep (B, H, T, Tp) uint8 with Tp a multiple of 128, scale = 1 / (1 - p)): training
    forward with attention-probability dropout (tavsr_relpos_attn_fwd_dropout, fp32 only)."""
    _chk2d(qkv, "qkv", None)
    _chk2d(pos, "pos", qkv.dtype)
    if qkv.shape != (B * T, 3 * H * 64) or pos.shape != (2 * T - 1, H * 64) or u.numel() != H * 64 \
            or v.numel() != H * 64:
        raise _lib.TavsrError(
            f"relpos_attn: the attention kernel is built for head width d_k = 64: expected qkv "
            f"({B * T}, {3 * H * 64}), pos ({2 * T - 1}, {H * 64}), u / v of {H * 64} values, got "
            f"{tuple(qkv.shape)}, {tuple(pos.shape)}, {u.numel()}, {v.numel()}")
    if out is None:
        out = torch.empty((B * T, H * 64), device=qkv.device, dtype=qkv.dtype)
    if drop is not None:
        keep, scale = drop
        if _is_bf16(qkv) or keep.dtype != torch.uint8 or keep.shape[:3] != (B, H, T) \
                or not keep.is_contiguous():
            raise _lib.TavsrError("relpos_attn: dropout needs fp32 operands and a contiguous uint8 "
                                  f"keep mask (B, H, T, Tp); got {keep.dtype} {tuple(keep.shape)}")
        check(_lib.load().tavsr_relpos_attn_fwd_dropout(
            qkv.data_ptr(), qkv.stride(0), pos.data_ptr(), pos.stride(0), u.data_ptr(), v.data_ptr(),
            _p(lens), out.data_ptr(), out.stride(0), B, T, H, _p(lse), keep.data_ptr(),
            keep.shape[3], float(scale), _stream()), "tavsr_relpos_attn_fwd_dropout")
        return out
    dt = (DT_BF16 | DT_OUT_BF16) if _is_bf16(qkv) else DT_TF32
    check(_lib.load().tavsr_relpos_attn_fwd(qkv.data_ptr(), qkv.stride(0), pos.data_ptr(),
                                            pos.stride(0), u.data_ptr(), v.data_ptr(), _p(lens),
                                            out.data_ptr(), out.stride(0), B, T, H,
                                            int(round_out), dt, _p(lse), _stream()),
          "tavsr_relpos_attn_fwd")
    return out


@_profiled
def csgu(h: torch.Tensor, norm_g: torch.Tensor, norm_b: torch.Tensor, conv_w: torch.Tensor,
         conv_b: torch.Tensor, B: int, T: int, eps: float = 1e-12, round_out: bool = True,
         out: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = r * (dwconv31(LN(g)) + b) for h = [r | g] (tavsr_csgu_fwd); h / out fp32 or bf16."""
    _chk2d(h, "h", None)
    Ch = h.shape[1] // 2
    if out is None:
        out = torch.empty((B * T, Ch), device=h.device, dtype=h.dtype)
    if stats is None:
        stats = torch.empty((B * T, 2), device=h.device, dtype=torch.float32)
    ksize = conv_w.shape[-1]
    dt = (DT_BF16 | DT_OUT_BF16) if _is_bf16(h) else DT_TF32
    check(_lib.load().tavsr_csgu_fwd(h.data_ptr(), h.stride(0), norm_g.data_ptr(),
                                     norm_b.data_ptr(), conv_w.data_ptr(), conv_b.data_ptr(),
                                     out.data_ptr(), out.stride(0), stats.data_ptr(), B, T, Ch,
                                     ksize, eps, int(round_out), dt, _stream()), "tavsr_csgu_fwd")
    return out


@_profiled
def merge_weights2(dots1: torch.Tensor, np1: int, dots2: torch.Tensor, np2: int,
                   lens1: Optional[torch.Tensor], lens2: Optional[torch.Tensor], pool_b1: float,
                   pool_b2: float, wproj_b1: float, wproj_b2: float, size: int, B: int, T: int):
    """learned_ave weights from partial row dots, one length array per branch
    (tavsr_merge_learned_ave_weights2)."""
    w1 = torch.empty((B,), device=dots1.device, dtype=torch.float32)
    w2 = torch.empty((B,), device=dots1.device, dtype=torch.float32)
    check(_lib.load().tavsr_merge_learned_ave_weights2(
        dots1.data_ptr(), np1, dots2.data_ptr(), np2, _p(lens1), _p(lens2), pool_b1, pool_b2,
        wproj_b1, wproj_b2, 1.0 / math.sqrt(size), w1.data_ptr(), w2.data_ptr(), B, T, _stream()),
        "tavsr_merge_learned_ave_weights2")
    return w1, w2


@_profiled
def scale_add_rows(a: torch.Tensor, b: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor,
                   rows_per_seg: int, out: Optional[torch.Tensor] = None,
                   out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """out[m] = w1[m // rows_per_seg] * a[m] + w2[m // rows_per_seg] * b[m] (tavsr_scale_add_rows);
    fp32 inputs, fp32 or bf16 output."""
    _chk2d(a, "a")
    _chk2d(b, "b")
    M, D = a.shape
    if out is None:
        out = torch.empty((M, D), device=a.device, dtype=out_dtype or torch.float32)
    check(_lib.load().tavsr_scale_add_rows(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0),
                                           w1.data_ptr(), w2.data_ptr(), rows_per_seg,
                                           out.data_ptr(), out.stride(0), M, D,
                                           DT_OUT_BF16 if _is_bf16(out) else 0, _stream()),
          "tavsr_scale_add_rows")
    return out


@_profiled
def split_tf32(x: torch.Tensor, kind: str) -> torch.Tensor:
    """3xTF32 operand triple (tavsr_split_tf32): kind "x" -> [hi | hi | lo], "w" -> [hi | lo | hi],
    (M, K) fp32 -> (M, 3K) fp32."""
    _chk2d(x, "x")
    M, K = x.shape
    out = torch.empty((M, 3 * K), device=x.device, dtype=torch.float32)
    check(_lib.load().tavsr_split_tf32(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), M, K,
                                       0 if kind == "x" else 1, _stream()), "tavsr_split_tf32")
    return out


_CAST_SCALARS = {}


def _cast_scalars(device):
    """Device scalars (1, 0) used as scale_add_rows weights."""
    sc = _CAST_SCALARS.get(device)
    if sc is None:
        sc = _CAST_SCALARS[device] = (torch.ones(1, device=device), torch.zeros(1, device=device))
    return sc


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 (M, D) -> bf16 copy on our own kernel (scale_add_rows with weights (1, 0)): the operand
    form of tensors that enter the bf16 path from outside (input features, pos_emb)."""
    if x.dtype == torch.bfloat16:
        return x
    sc = _cast_scalars(x.device)
    return scale_add_rows(x, x, sc[0], sc[1], max(1, x.shape[0]), out_dtype=torch.bfloat16)


@_profiled
def merge_weights(dots1: torch.Tensor, dots2: torch.Tensor, lens: Optional[torch.Tensor],
                  pool_b1: float, pool_b2: float, wproj_b1: float, wproj_b2: float, size: int,
                  B: int, T: int, w1: Optional[torch.Tensor] = None,
                  w2: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    if w1 is None:
        w1 = torch.empty((B,), device=dots1.device, dtype=torch.float32)
    if w2 is None:
        w2 = torch.empty((B,), device=dots1.device, dtype=torch.float32)
    check(_lib.load().tavsr_merge_learned_ave_weights(
        dots1.data_ptr(), dots2.data_ptr(), _p(lens), pool_b1, pool_b2, wproj_b1, wproj_b2,
        1.0 / math.sqrt(size), w1.data_ptr(), w2.data_ptr(), B, T, _stream()),
        "tavsr_merge_learned_ave_weights")
    return w1, w2


@_profiled
def merge_scores(a1: torch.Tensor, a2: torch.Tensor, va1: torch.Tensor, vb1: torch.Tensor,
                 va2: torch.Tensor, vb2: torch.Tensor, lens: Optional[torch.Tensor], pool_b1: float,
                 pool_b2: float, wproj_b1: float, wproj_b2: float, size: int, B: int, T: int):
    """learned_ave merge weights straight from the branch activations in ONE launch
    (tavsr_merge_scores: row dots + masked softmax pooling + 2-way softmax, a cluster of 4 CTAs per
    utterance).  a1 (B*T, 256), a2 (B*T, 1024), fp32 or bf16 (both the same)."""
    _chk2d(a1, "a1", None)
    _chk2d(a2, "a2", a1.dtype)
    w1 = torch.empty((B,), device=a1.device, dtype=torch.float32)
    w2 = torch.empty((B,), device=a1.device, dtype=torch.float32)
    check(_lib.load().tavsr_merge_scores(
        a1.data_ptr(), a1.stride(0), a1.shape[1], a2.data_ptr(), a2.stride(0), a2.shape[1],
        va1.data_ptr(), vb1.data_ptr(), va2.data_ptr(), vb2.data_ptr(), _p(lens), pool_b1, pool_b2,
        wproj_b1, wproj_b2, 1.0 / math.sqrt(size), w1.data_ptr(), w2.data_ptr(), B, T,
        DT_BF16 if _is_bf16(a1) else DT_TF32, _stream()), "tavsr_merge_scores")
    return w1, w2


@_profiled
def merge_weights_dev(dots1: torch.Tensor, dots2: torch.Tensor, lens: Optional[torch.Tensor],
                      scal: torch.Tensor, size: int, B: int, T: int, lens2: Optional[torch.Tensor] = None):
    """merge_weights with the four biases (pool_b1, pool_b2, wproj_b1, wproj_b2) read from the device
    tensor `scal` (training: parameters change every step, no host read-back); `lens2`: branch 2
    masked by its own lengths (audio-visual fusion)."""
    w1 = torch.empty((B,), device=dots1.device, dtype=torch.float32)
    w2 = torch.empty((B,), device=dots1.device, dtype=torch.float32)
    check(_lib.load().tavsr_merge_learned_ave_weights_dev(
        dots1.data_ptr(), dots2.data_ptr(), _p(lens), _p(lens2), scal.data_ptr(), 1.0 / math.sqrt(size),
        w1.data_ptr(), w2.data_ptr(), B, T, _stream()), "tavsr_merge_learned_ave_weights_dev")
    return w1, w2


@_profiled
def row_dots(a1: torch.Tensor, va1: torch.Tensor, vb1: torch.Tensor,
             a2: Optional[torch.Tensor] = None, va2: Optional[torch.Tensor] = None,
             vb2: Optional[torch.Tensor] = None):
    """(a1 @ [va1 vb1], a2 @ [va2 vb2]) as (M,2) tensors in one launch (tavsr_row_dots); a1 / a2
    fp32 or bf16 (both the same), vectors fp32."""
    _chk2d(a1, "a1", None)
    if a2 is not None:
        _chk2d(a2, "a2", a1.dtype)
    M = a1.shape[0]
    o1 = torch.empty((M, 2), device=a1.device, dtype=torch.float32)
    o2 = torch.empty((M, 2), device=a1.device, dtype=torch.float32) if a2 is not None else None
    check(_lib.load().tavsr_row_dots(
        a1.data_ptr(), a1.stride(0), a1.shape[1], va1.data_ptr(), vb1.data_ptr(), o1.data_ptr(),
        _p(a2), a2.stride(0) if a2 is not None else 0, a2.shape[1] if a2 is not None else 0,
        _p(va2), _p(vb2), _p(o2), M, DT_BF16 if _is_bf16(a1) else DT_TF32, _stream()), "tavsr_row_dots")
    return o1, o2


@_profiled
def conv2d_sub_im2col(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor,
                      out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """im2col operand of the second Conv2dSubsampling convolution with conv1 + ReLU evaluated on
    the fly: x (B, Tin, F) -> A (B*T2*F2, 9*C) (tavsr_conv2d_sub_im2col)."""
    if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
        raise _lib.TavsrError("conv2d_sub_im2col: x must be a contiguous fp32 CUDA tensor")
    B, Tin, F = x.shape
    C = w1.shape[0]
    T2, F2 = ((Tin - 1) // 2 - 1) // 2, ((F - 1) // 2 - 1) // 2
    A = torch.empty((B * T2 * F2, 9 * C), device=x.device, dtype=out_dtype)
    check(_lib.load().tavsr_conv2d_sub_im2col(x.data_ptr(), B, Tin, F, w1.data_ptr(), b1.data_ptr(),
                                              C, A.data_ptr(), DT_OUT_BF16 if _is_bf16(A) else 0,
                                              _stream()), "tavsr_conv2d_sub_im2col")
    return A


@_profiled
def ctc_head(hs: torch.Tensor, w: torch.Tensor, b: torch.Tensor, want_logp: bool = True,
             want_prob: bool = False, want_argmax: bool = False, want_logits: bool = False):
    """(logp, prob, argmax[, logits]) of ctc_lo(hs) over the vocabulary, fp32 FMA
    (tavsr_ctc_head)."""
    _chk2d(hs, "hs")
    M, D = hs.shape
    V = w.shape[0]
    logp = torch.empty((M, V), device=hs.device, dtype=torch.float32) if want_logp else None
    prob = torch.empty((M, V), device=hs.device, dtype=torch.float32) if want_prob else None
    amax = torch.empty((M,), device=hs.device, dtype=torch.int64) if want_argmax else None
    logits = torch.empty((M, V), device=hs.device, dtype=torch.float32) if want_logits else None
    check(_lib.load().tavsr_ctc_head(hs.data_ptr(), hs.stride(0), w.data_ptr(), b.data_ptr(),
                                     _p(logits), _p(logp), _p(prob), _p(amax), M, D, V, _stream()),
          "tavsr_ctc_head")
    if want_logits:
        return logp, prob, amax, logits
    return logp, prob, amax


@_profiled
def vocab_residual(x: torch.Tensor, p: torch.Tensor, w: torch.Tensor, b: torch.Tensor,
                   ln: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, eps: float = 1e-12,
                   ln_dtype: torch.dtype = torch.float32):
    """out = x + p @ w.T + b  (p (M,V) posteriors, w (D,V)); with `ln` also LayerNorm(out)
    (tavsr_vocab_residual).  Returns (out, xn-or-None)."""
    _chk2d(x, "x")
    _chk2d(p, "p")
    M, D = x.shape
    V = p.shape[1]
    if w.shape != (D, V) or not w.is_contiguous() or not p.is_contiguous():
        raise ValueError("vocab_residual: w must be a contiguous (D,V) matrix, p contiguous (M,V)")
    if V > 64:
        w = w.t().contiguous()   # the large-vocabulary kernel streams W^T (V, D) rows from L2
    out = torch.empty((M, D), device=x.device, dtype=torch.float32)
    xn = torch.empty((M, D), device=x.device, dtype=ln_dtype) if ln is not None else None
    check(_lib.load().tavsr_vocab_residual(
        x.data_ptr(), x.stride(0), p.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(),
        out.stride(0), _p(ln[0]) if ln else None, _p(ln[1]) if ln else None, eps, _p(xn),
        xn.stride(0) if xn is not None else 0, M, D, V, DT_LNA_BF16 if _is_bf16(xn) else 0,
        _stream()), "tavsr_vocab_residual")
    return out, xn


@_profiled
def ctc_loss(logp: torch.Tensor, targets: torch.Tensor, hlens: torch.Tensor, tlens: torch.Tensor,
             want_grad: bool = False, gscale: float = 1.0, zero_infinity: bool = True):
    """Per-utterance NLL (and optionally d/dlogits) from (B,T,V) log-probs (tavsr_ctc_loss)."""
    B, T, V = logp.shape
    Lmax = targets.shape[1]
    nll = torch.empty((B,), device=logp.device, dtype=torch.float32)
    grad = ws = None
    lib = _lib.load()
    if want_grad:
        grad = torch.empty((B, T, V), device=logp.device, dtype=torch.float32)
        ws = torch.empty((lib.tavsr_ctc_workspace_bytes(B, T, Lmax) // 4,), device=logp.device,
                         dtype=torch.float32)
    check(lib.tavsr_ctc_loss(logp.data_ptr(), targets.data_ptr(), targets.stride(0),
                             hlens.data_ptr(), tlens.data_ptr(), nll.data_ptr(), _p(grad), gscale,
                             _p(ws), B, T, V, Lmax, int(zero_infinity), _stream()),
          "tavsr_ctc_loss")
    return nll, grad


@_profiled
def ctc_head_bwd(dlogits: torch.Tensor, hs: torch.Tensor, w: torch.Tensor,
                 row_scale: Optional[torch.Tensor] = None, rows_per_seg: int = 0):
    """Backward of the CTC head: (dhs (M,D), dw (V,D), db (V,)) from dlogits (M,V), optionally
    scaled per utterance (tavsr_ctc_head_bwd)."""
    _chk2d(hs, "hs")
    M, D = hs.shape
    V = w.shape[0]
    dl = dlogits.reshape(M, V)
    if not dl.is_contiguous() or not w.is_contiguous():
        raise _lib.TavsrError("ctc_head_bwd: dlogits and w must be contiguous")
    lib = _lib.load()
    dhs = torch.empty((M, D), device=hs.device, dtype=torch.float32)
    dw = torch.empty((V, D), device=hs.device, dtype=torch.float32)
    db = torch.empty((V,), device=hs.device, dtype=torch.float32)
    ws = torch.empty((int(lib.tavsr_ctc_head_bwd_workspace_bytes(M)) // 4,), device=hs.device,
                     dtype=torch.float32)
    check(lib.tavsr_ctc_head_bwd(dl.data_ptr(), _p(row_scale), rows_per_seg, hs.data_ptr(),
                                 hs.stride(0), w.data_ptr(), dhs.data_ptr(), dhs.stride(0),
                                 dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel() * 4, M, D, V,
                                 _stream()), "tavsr_ctc_head_bwd")
    return dhs, dw, db


@_profiled
def ctc_greedy(amax: torch.Tensor, lens: Optional[torch.Tensor], blank: int = 0):
    B, T = amax.shape
    tokens = torch.empty((B, T), device=amax.device, dtype=torch.int64)
    ntok = torch.empty((B,), device=amax.device, dtype=torch.int32)
    check(_lib.load().tavsr_ctc_greedy(amax.data_ptr(), _p(lens), tokens.data_ptr(),
                                       ntok.data_ptr(), B, T, blank, _stream()),
          "tavsr_ctc_greedy")
    return tokens, ntok


@_profiled
def ctc_prefix_score(logp: torch.Tensor, r_prev: torch.Tensor, last: torch.Tensor,
                     plen: torch.Tensor, psi_prev: torch.Tensor, Tvalid: int, blank: int, eos: int):
    T, V = logp.shape
    nhyp = r_prev.shape[0]
    r_new = torch.empty((nhyp, T, V, 2), device=logp.device, dtype=torch.float32)
    score = torch.empty((nhyp, V), device=logp.device, dtype=torch.float32)
    check(_lib.load().tavsr_ctc_prefix_score(logp.data_ptr(), r_prev.data_ptr(), last.data_ptr(),
                                             plen.data_ptr(), psi_prev.data_ptr(),
                                             r_new.data_ptr(), score.data_ptr(), T, Tvalid, V, nhyp,
                                             blank, eos, _stream()), "tavsr_ctc_prefix_score")
    return r_new, score
