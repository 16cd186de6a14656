"""Tensor-level wrappers over the backward kernels of the C ABI (include/tavsr.h, csrc/backward.cu,
csrc/attention_bwd.cu; GPU-tested in tests/test_backward_gpu.py).  training.py composes them into
the autograd nodes of the encoder; the arithmetic is fixed by oracle/bwd_formulas.py."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check
from .ops import _chk2d, _p, _stream


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty((max(1, int(nbytes)) + 3) // 4, dtype=torch.float32, device=device)


def transpose_2d(x: torch.Tensor, pad: bool = False) -> torch.Tensor:
    """out[c][r] = x[r][c].  With `pad` the output rows are padded with zeros to a multiple of 4
    columns (the operand form of a wgrad product, whose reduction axis is the row count of x)."""
    _chk2d(x, "x")
    R, C = x.shape
    Rp = (R + 3) // 4 * 4 if pad else R
    out = torch.empty((C, Rp), device=x.device, dtype=torch.float32)
    check(_lib.load().tavsr_transpose_2d(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), R, C,
                                         _stream()), "tavsr_transpose_2d")
    return out


def act_fwd(z: torch.Tensor, act: int, mask: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """h = act(z) (* mask): the training forward keeps z and evaluates h separately."""
    _chk2d(z, "z")
    M, C = z.shape
    if out is None:
        out = torch.empty((M, C), device=z.device, dtype=torch.float32)
    check(_lib.load().tavsr_act_fwd(z.data_ptr(), z.stride(0), _p(mask),
                                    mask.stride(0) if mask is not None else 0, out.data_ptr(),
                                    out.stride(0), M, C, act, _stream()), "tavsr_act_fwd")
    return out


def act_fwd_t(z: torch.Tensor, act: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(act(z) * mask)^T as a (C, ceil4(M)) wgrad operand in one pass (tavsr_act_fwd_t)."""
    _chk2d(z, "z")
    M, C = z.shape
    out = torch.empty((C, (M + 3) // 4 * 4), device=z.device, dtype=torch.float32)
    check(_lib.load().tavsr_act_fwd_t(z.data_ptr(), z.stride(0), _p(mask),
                                      mask.stride(0) if mask is not None else 0, out.data_ptr(),
                                      out.stride(0), M, C, act, _stream()), "tavsr_act_fwd_t")
    return out


def act_bwd_t(z: torch.Tensor, dh: torch.Tensor, act: int):
    """dz = dh * act'(z) row-major plus its transpose (C, ceil4(M)) in one pass (tavsr_act_bwd_t)."""
    _chk2d(z, "z")
    _chk2d(dh, "dh")
    M, C = z.shape
    dz = torch.empty((M, C), device=z.device, dtype=torch.float32)
    dzT = torch.empty((C, (M + 3) // 4 * 4), device=z.device, dtype=torch.float32)
    check(_lib.load().tavsr_act_bwd_t(z.data_ptr(), z.stride(0), dh.data_ptr(), dh.stride(0),
                                      dz.data_ptr(), dz.stride(0), dzT.data_ptr(), dzT.stride(0), M, C,
                                      act, _stream()), "tavsr_act_bwd_t")
    return dz, dzT


def gemm_wgrad(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """out (M, N) = a (M, K) . b (N, K)^T with the reduction axis split over the machine
    (tavsr_gemm_wgrad): the weight-gradient product, K = frames of the batch."""
    _chk2d(a, "a")
    _chk2d(b, "b")
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K:
        raise _lib.TavsrError(f"gemm_wgrad: reduction axes differ ({K} vs {b.shape[1]})")
    lib = _lib.load()
    out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    nbytes = lib.tavsr_gemm_wgrad_workspace_bytes(M, N, K)
    ws = _ws(nbytes, a.device) if nbytes else None
    check(lib.tavsr_gemm_wgrad(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                               out.stride(0), M, N, K, _p(ws), nbytes, _stream()), "tavsr_gemm_wgrad")
    return out


def linear_bwd(x_in: Optional[torch.Tensor], w: torch.Tensor, dy: torch.Tensor, need_dx: bool = True,
               need_dw: bool = True, wT: Optional[torch.Tensor] = None,
               xT: Optional[torch.Tensor] = None, dyT: Optional[torch.Tensor] = None):
    """Backward of y = x_in W^T + b on the tcgen05 GEMM (TF32 operands, fp32 accumulate):
         dx = dy W            = gemm(dy, (W^T) as the (K, N) weight operand)
         dW = dy^T x_in       = gemm(dy^T (N, M), x_in^T (K, M)): reduction over the M rows
         db = column sums of dy
    Both products want K-major operands, hence the transposed copies (a 32 x 32 tile transpose
    kernel; the bf16 path can read MN-major operands directly and will not need them).
    `xT` / `dyT`: operands already transposed by a fused producer (act_fwd_t / act_bwd_t).
    Returns (dx or None, dW or None, db or None)."""
    from . import ops
    dx = dw = db = None
    if need_dx:
        if wT is None:
            wT = transpose_2d(w.contiguous() if w.stride(1) != 1 else w)
        dx = ops.gemm_bias_act(dy, wT, None)
    if need_dw:
        if dyT is None:
            dyT = transpose_2d(dy, pad=True)
        if xT is None:
            xT = transpose_2d(x_in, pad=True)
        dw = gemm_wgrad(dyT, xT)
        db = col_sums(dy)
    return dx, dw, db


def relpos_attn_bwd(qkv: torch.Tensor, pos: torch.Tensor, u: torch.Tensor, v: torch.Tensor,
                    lens: Optional[torch.Tensor], ctx: torch.Tensor, dctx: torch.Tensor,
                    lse: torch.Tensor, B: int, T: int, H: int, drop: Optional[tuple] = None):
    """Backward of ops.relpos_attn (tavsr_relpos_attn_bwd); `drop` = the forward's (keep, scale).  Returns (dqkv (B*T, 3*H*64), dpos
    (2T-1, H*64), du (H*64,), dv (H*64,))."""
    from . import ops
    for t, n in ((qkv, "qkv"), (pos, "pos"), (ctx, "ctx"), (dctx, "dctx")):
        _chk2d(t, n)
    M, HD = B * T, H * 64
    dev = qkv.device
    R = 2 * T - 1
    dqkv = torch.zeros((M, 3 * HD), device=dev, dtype=torch.float32)     # the q block is accumulated
    acc = torch.zeros(B * R * HD + 2 * HD, device=dev, dtype=torch.float32)   # one memset
    dpos_b, du, dv = acc[:B * R * HD], acc[B * R * HD:B * R * HD + HD], acc[B * R * HD + HD:]
    check(_lib.load().tavsr_relpos_attn_bwd(
        qkv.data_ptr(), qkv.stride(0), pos.data_ptr(), pos.stride(0), u.data_ptr(), v.data_ptr(),
        _p(lens), ctx.data_ptr(), ctx.stride(0), dctx.data_ptr(), dctx.stride(0), lse.data_ptr(),
        dqkv.data_ptr(), dqkv.stride(0), du.data_ptr(), dv.data_ptr(), dpos_b.data_ptr(),
        drop[0].data_ptr() if drop is not None else None, drop[0].shape[3] if drop is not None else 0,
        float(drop[1]) if drop is not None else 1.0, B, T, H, _stream()), "tavsr_relpos_attn_bwd")
    dpos = col_sums(dpos_b.view(B, R * HD)).view(R, HD)     # sum of the per-utterance slabs
    return dqkv, dpos, du, dv


def col_sums(a: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk2d(a, "a")
    R, C = a.shape
    lib = _lib.load()
    out = torch.empty((C,), device=a.device, dtype=torch.float32)
    ws = _ws(lib.tavsr_col_sums_workspace_bytes(R, C), a.device)
    check(lib.tavsr_col_sums(a.data_ptr(), a.stride(0), _p(b), b.stride(0) if b is not None else 0,
                             out.data_ptr(), ws.data_ptr(), ws.numel() * 4, R, C, _stream()),
          "tavsr_col_sums")
    return out


def act_bwd(z: torch.Tensor, dh: torch.Tensor, act: int) -> torch.Tensor:
    _chk2d(z, "z")
    _chk2d(dh, "dh")
    M, C = z.shape
    dz = torch.empty((M, C), device=z.device, dtype=torch.float32)
    check(_lib.load().tavsr_act_bwd(z.data_ptr(), z.stride(0), dh.data_ptr(), dh.stride(0),
                                    dz.data_ptr(), dz.stride(0), M, C, act, _stream()), "tavsr_act_bwd")
    return dz


def layernorm_bwd(x: torch.Tensor, gamma: torch.Tensor, dy: torch.Tensor, eps: float = 1e-12,
                  dres: Optional[torch.Tensor] = None, dx: Optional[torch.Tensor] = None):
    """Returns (dx, dgamma, dbeta); x / dy / dx may be strided row views (e.g. the gate half of h)."""
    _chk2d(x, "x")
    _chk2d(dy, "dy")
    M, D = x.shape
    lib = _lib.load()
    if dx is None:
        dx = torch.empty((M, D), device=x.device, dtype=torch.float32)
    gb = torch.empty((2, D), device=x.device, dtype=torch.float32)
    ws = _ws(lib.tavsr_layernorm_bwd_workspace_bytes(M, D), x.device)
    check(lib.tavsr_layernorm_bwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), dy.data_ptr(),
                                  dy.stride(0), _p(dres), dres.stride(0) if dres is not None else 0,
                                  dx.data_ptr(), dx.stride(0), gb[0].data_ptr(), gb[1].data_ptr(),
                                  ws.data_ptr(), ws.numel() * 4, M, D, eps, _stream()),
          "tavsr_layernorm_bwd")
    return dx, gb[0], gb[1]


def csgu_bwd(h: torch.Tensor, norm_g: torch.Tensor, norm_b: torch.Tensor, conv_w: torch.Tensor,
             conv_b: torch.Tensor, stats: torch.Tensor, du: torch.Tensor, B: int, T: int,
             eps: float = 1e-12):
    """Full CSGU backward: (dh (B*T, 2Ch), dnorm_g, dnorm_b, dconv_w (Ch,31), dconv_b).  `stats` is
    the forward's (mean, rstd) per frame (tavsr_csgu_fwd's scratch)."""
    _chk2d(h, "h")
    _chk2d(du, "du")
    Ch = h.shape[1] // 2
    lib = _lib.load()
    dh = torch.empty_like(h)
    dn = torch.empty((B * T, Ch), device=h.device, dtype=torch.float32)
    dcw = torch.empty((Ch, conv_w.shape[-1]), device=h.device, dtype=torch.float32)
    dcb = torch.empty((Ch,), device=h.device, dtype=torch.float32)
    ws = _ws(lib.tavsr_csgu_bwd_workspace_bytes(B, T, Ch), h.device)
    check(lib.tavsr_csgu_conv_bwd(h.data_ptr(), h.stride(0), norm_g.data_ptr(), norm_b.data_ptr(),
                                  conv_w.data_ptr(), conv_b.data_ptr(), stats.data_ptr(), du.data_ptr(),
                                  du.stride(0), dh.data_ptr(), dh.stride(0), dn.data_ptr(), dn.stride(0),
                                  dcw.data_ptr(), dcb.data_ptr(), ws.data_ptr(), ws.numel() * 4, B, T, Ch,
                                  conv_w.shape[-1], _stream()), "tavsr_csgu_conv_bwd")
    _, dng, dnb = layernorm_bwd(h[:, Ch:], norm_g, dn, eps=eps, dx=dh[:, Ch:])
    return dh, dng, dnb, dcw, dcb


def softmax_bwd(p: torch.Tensor, dp: torch.Tensor) -> torch.Tensor:
    """dlogits = p * (dp - sum_v p dp) over the last axis of contiguous (M, V) tensors
    (tavsr_softmax_bwd)."""
    if p.shape != dp.shape or not p.is_contiguous() or not dp.is_contiguous() or p.dtype != torch.float32:
        raise _lib.TavsrError("softmax_bwd: p and dp must be contiguous fp32 tensors of one shape")
    M, V = p.shape
    out = torch.empty_like(p)
    check(_lib.load().tavsr_softmax_bwd(p.data_ptr(), dp.data_ptr(), out.data_ptr(), M, V, _stream()),
          "tavsr_softmax_bwd")
    return out


def conv2d_sub_bwd(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, dA: torch.Tensor):
    """(d conv1.weight (C, 9), d conv1.bias (C,)) of the Conv2dSubsampling front end from the
    gradient dA of its im2col operand (tavsr_conv2d_sub_bwd); x (B, Tin, F) fp32, w1 (C, 9)."""
    B, Tin, F = x.shape
    C = w1.shape[0]
    lib = _lib.load()
    nbytes = lib.tavsr_conv2d_sub_bwd_workspace_bytes(B, Tin, C)
    ws = _ws(nbytes, x.device)
    grads = torch.empty((C, 10), device=x.device, dtype=torch.float32)
    check(lib.tavsr_conv2d_sub_bwd(x.data_ptr(), B, Tin, F, w1.data_ptr(), b1.data_ptr(), C,
                                   dA.data_ptr(), grads.data_ptr(), ws.data_ptr(), nbytes, _stream()),
          "tavsr_conv2d_sub_bwd")
    return grads[:, :9].contiguous(), grads[:, 9].contiguous()


def merge_learned_ave_bwd(x1: torch.Tensor, x2: torch.Tensor, dm: torch.Tensor, lens: torch.Tensor,
                          a1: torch.Tensor, b1: torch.Tensor, a2: torch.Tensor, b2: torch.Tensor,
                          scal: torch.Tensor, B: int, T: int, lens2: Optional[torch.Tensor] = None):
    """Returns (dx1, dx2, grads (1028,)): see tavsr_merge_learned_ave_bwd in include/tavsr.h.
    scal: device tensor [c1, e1, c2, e2] (pooling_proj / weight_proj biases); lens2: branch 2's own
    lengths (audio-visual fusion), default = lens."""
    _chk2d(x1, "x1")
    _chk2d(x2, "x2")
    _chk2d(dm, "dm")
    M, D = x1.shape
    lib = _lib.load()
    dx1 = torch.empty((M, D), device=x1.device, dtype=torch.float32)
    dx2 = torch.empty((M, D), device=x1.device, dtype=torch.float32)
    grads = torch.empty((4 * D + 4,), device=x1.device, dtype=torch.float32)
    ws = _ws(lib.tavsr_merge_learned_ave_bwd_workspace_bytes(B), x1.device)
    check(lib.tavsr_merge_learned_ave_bwd(
        x1.data_ptr(), x1.stride(0), x2.data_ptr(), x2.stride(0), dm.data_ptr(), dm.stride(0),
        _p(lens), _p(lens2), a1.data_ptr(), b1.data_ptr(), a2.data_ptr(), b2.data_ptr(), scal.data_ptr(),
        dx1.data_ptr(), dx1.stride(0), dx2.data_ptr(), dx2.stride(0), grads.data_ptr(), ws.data_ptr(),
        ws.numel() * 4, B, T, D, _stream()), "tavsr_merge_learned_ave_bwd")
    return dx1, dx2, grads
