"""Kernel sequencing for one Branchformer block on the B200 path.

The layer modules own the parameters; this file turns them into launch sequences of the C-ABI
kernels (ops.py).  All activations are 2-D row-major (B*T, C) tensors.  Compute modes
(`set_compute_dtype`, env TAVSR_DTYPE):
  "tf32"   (default) fp32 storage everywhere, operands rounded to TF32 by TMA on load;
  "tf32x3" as tf32, but the dense projections run as three-term split TF32 products
           (a_hi b_hi + a_hi b_lo + a_lo b_hi over a tripled reduction axis): fp32-class accuracy
           on the TF32 pipe, the parity fallback for checkpoints whose statistics eat the TF32 margin;
  "bf16"   every tensor-core operand is stored as bf16 by the kernel that produces it (LayerNorm
           outputs, qkv, ctx, the GELU'd cgMLP hidden, the CSGU output) and weights are converted
           once per parameter version; the residual stream, LayerNorm statistics, softmax, the
           CSGU arithmetic, the encoder output and the whole CTC scorer stay fp32.
Accumulation is fp32 in every mode.

Fusion map of one two-branch `learned_ave` block (reference encoder_layer.py:153-321):

  gemm_bias_act   h   = swish(LN_ffmac(x) W1^T + b1)                           [:193-194]
  gemm_rowln      x   = x + .5 (h W2^T + b2);  xa = LN_mha(x), xm = LN_mlp(x)   [:194,202,216]
  gemm_bias_act   qkv = xa Wqkv^T + bqkv                                        [:208]
  relpos_attn     ctx = softmax(((q+u)k^T + shift((q+v)p^T))/8) v               [:208]
  gemm_rowln      x1  = ctx Wo^T + bo;  (s1,z1) = x1.(pool1, wproj1)            [:208,243,258]
  gemm_bias_act   g   = gelu(xm Wc1^T + bc1)                                    [:220]
  csgu            u   = g_r * (dwconv31(LN(g_g)) + cb)                          [:220]
  gemm_rowln      x2  = u Wc2^T + bc2;  (s2,z2) = x2.(pool2, wproj2)            [:220,262,277]
  merge_weights   (w1,w2) per utterance                                         [:245-289]
  gemm_rowln      x   = x + (w1 x1 + w2 x2) Wm^T + bm;  xf = LN_ff(x)           [:291-293,313]
  gemm_bias_act   h   = swish(xf W1^T + b1)                                     [:314]
  gemm_rowln      y   = LN_final(x + .5 (h W2^T + b2));  yn = LN_next(y)        [:314,316]

`LN_next` is the next block's norm_ff_macaron (or the encoder's after_norm), so no stand-alone
LayerNorm kernel runs between blocks.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops

import os

FFN_FUSED = os.environ.get("TAVSR_FFN_FUSED", "1") != "0"
# fold attn.linear_out and cgmlp.channel_proj2 into merge_proj (two-branch learned_ave / fixed_ave
# blocks): the merge GEMM then reads the attention context and the gated activations directly
FOLD_MERGE = os.environ.get("TAVSR_FOLD_MERGE", "1") != "0"

# learned_ave merge weights by the fused cluster kernel (row dots + pooling + softmax, one launch)
# instead of row_dots + merge_weights (TAVSR_FUSE_SCORES=0: the two-kernel sequence)
FUSE_SCORES = os.environ.get("TAVSR_FUSE_SCORES", "1") != "0"

# run the attention and cgMLP branches of a two-branch block on two streams.  MEASURED (C2, graph
# replay): 3.786 vs 3.789 ms per step - every kernel of the block already fills the GPU (persistent
# GEMMs, 1 CTA/SM attention), so the branches serialise anyway; opt-in.
BRANCH_FORK = os.environ.get("TAVSR_BRANCH_FORK", "0") != "0"
_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class branch_fork:
    """`with branch_fork(dev) as side:` — inside, `with side:` runs launches on a side stream that
    was made to wait for everything enqueued on the current stream so far; leaving the outer block
    makes the current stream wait for the side stream (fork / join; under CUDA-graph capture the
    two chains become parallel graph branches).  Tensors produced on the side stream are only
    consumed after the join, and every later fork waits for the consumers first, so the caching
    allocator's per-stream reuse stays ordered."""

    def __init__(self, device, enabled=None):
        self.device = device
        self.side = None
        self.enabled = BRANCH_FORK if enabled is None else enabled

    def __enter__(self):
        if not self.enabled:
            return _NullCtx()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        side = _SIDE_STREAMS.get(idx)
        if side is None:
            side = _SIDE_STREAMS[idx] = torch.cuda.Stream(device=self.device)
        self.side = side
        side.wait_stream(torch.cuda.current_stream(self.device))
        return torch.cuda.stream(side)

    def __exit__(self, *exc):
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        return False


# Conv2dSubsampling on cuDNN instead of the im2col + tcgen05 path (cross-check knob)
CUDNN_EMBED = os.environ.get("TAVSR_CUDNN_EMBED", "0") != "0"

_ACT = {"swish": ops.ACT_SWISH, "relu": ops.ACT_RELU, "gelu": ops.ACT_GELU}

_MODES = ("tf32", "tf32x3", "bf16")
_DTYPE = os.environ.get("TAVSR_DTYPE", "tf32")
if _DTYPE not in _MODES:
    raise ValueError(f"TAVSR_DTYPE={_DTYPE!r}: expected one of {_MODES}")


def set_compute_dtype(mode: str) -> None:
    """Select the compute mode of every B200 module in this process (see the module docstring).
    Captured CUDA graphs key on it, derived-weight caches hold one entry per mode."""
    global _DTYPE
    if mode not in _MODES:
        raise ValueError(f"compute dtype {mode!r}: expected one of {_MODES}")
    _DTYPE = mode


def compute_dtype() -> str:
    return _DTYPE


class use_compute_dtype:
    """`with use_compute_dtype("bf16"): ...` (tests, benches)."""

    def __init__(self, mode: str):
        self.mode = mode

    def __enter__(self):
        self.prev = _DTYPE
        set_compute_dtype(self.mode)
        return self

    def __exit__(self, *exc):
        set_compute_dtype(self.prev)
        return False


def is_bf16() -> bool:
    return _DTYPE == "bf16"


def act_dtype() -> torch.dtype:
    """Storage type of tensors that only feed tensor-core products."""
    return torch.bfloat16 if _DTYPE == "bf16" else torch.float32


def operand(x: torch.Tensor) -> torch.Tensor:
    """Activation entering the path from outside (features, pos_emb) in operand storage."""
    return ops.cast_bf16(x) if _DTYPE == "bf16" and x.dtype != torch.bfloat16 else x


def act_code(name: str) -> int:
    if name not in _ACT:
        raise ValueError(f"ffn_activation_type={name!r} is not built on the B200 path "
                         f"(supported: {sorted(_ACT)})")
    return _ACT[name]


class PackedCache:
    """Derived tensors (fused QKV weight, flattened conv taps, ...) keyed on the versions of their
    source parameters, so in-place updates (optimizer steps, load_state_dict) invalidate them."""

    def __init__(self):
        self._store: Dict[str, Tuple[tuple, object]] = {}

    @staticmethod
    def _version(t) -> int:
        try:
            return t._version
        except RuntimeError:  # inference tensors do not track versions
            return -1

    def clear(self) -> None:
        self._store.clear()

    def weight(self, key: str, w: torch.Tensor) -> torch.Tensor:
        """Operand form of a weight matrix in the current compute mode: the parameter itself
        (tf32: TMA rounds on load) or a bf16 copy made once per parameter version."""
        if _DTYPE != "bf16":
            return w
        return self.get(key + ":bf16", [w], lambda: w.detach().to(torch.bfloat16).contiguous())

    def get(self, key: str, sources, build):
        sig = tuple((s.data_ptr(), self._version(s), s.device) for s in sources)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[key] = (sig, val)
        return val


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("tailored_avsr_b200 runs on CUDA tensors only (there is no CPU "
                               "fallback); move the module and its inputs to a B200 device")


def require_inference(module: torch.nn.Module, *tensors) -> None:
    """For the pieces whose backward is not built (InterCTC conditioning): refuse a grad-mode call
    loudly instead of silently detaching the graph.  Every encoder / fusion / embed / CTC module
    routes grad-mode calls to training.py instead."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise NotImplementedError(
            f"tailored_avsr_b200: {type(module).__name__} has no CUDA backward yet; call it under "
            "torch.no_grad() / torch.inference_mode() (see README 'Training').")
    require_cuda(*tensors)


def lens_from_mask(mask: Optional[torch.Tensor], B: int, T: int, device) -> torch.Tensor:
    """(B,1,T) bool key mask -> int32 lengths.  The kernels implement prefix masks (what
    make_pad_mask produces, encoder.py:345)."""
    if mask is None:
        return torch.full((B,), T, dtype=torch.int32, device=device)
    return mask.reshape(B, -1).sum(dim=1).to(torch.int32)


def ffn_block(x, xn, ff, *, out_main, ln0=None, lnA=None, out_lnA=None, round_lnA=False,
              lnB=None, out_lnB=None, eps0=1e-12, alpha=0.5, cache: Optional[PackedCache] = None,
              key: str = "ffn"):
    """x + alpha * (W2 act(W1 xn + b1) + b2) with the trailing LayerNorms fused (x may be None: no
    residual).  One kernel with the hidden activation kept on chip when the shapes are the built
    ones (256 -> 2048 -> 256), otherwise (or with TAVSR_FFN_FUSED=0, or in tf32x3 mode) the
    two-GEMM sequence.  Output storage types follow the tensors passed in."""
    cache = cache if cache is not None else _ffn_cache(ff)
    w1 = cache.weight(key + ".w1", ff.w_1.weight)
    w2 = cache.weight(key + ".w2", ff.w_2.weight)
    if FFN_FUSED and _DTYPE != "tf32x3" and tuple(ff.w_1.weight.shape) == (2048, 256):
        ops.ffn_fused(xn, w1, ff.w_1.bias, w2, ff.w_2.bias,
                      act_code(ff.activation_type), residual=x, alpha=alpha, ln0=ln0, eps0=eps0,
                      out_main=out_main, lnA=lnA, out_lnA=out_lnA, round_lnA=round_lnA,
                      lnB=lnB, out_lnB=out_lnB)
        return
    h = linear(xn, ff.w_1.weight, ff.w_1.bias, cache, key + ".w1", act=act_code(ff.activation_type),
               out_dtype=act_dtype())
    linear_rowln(h, ff.w_2.weight, ff.w_2.bias, cache, key + ".w2", residual=x, alpha=alpha, ln0=ln0,
                 eps0=eps0, out_main=out_main, lnA=lnA, out_lnA=out_lnA, round_lnA=round_lnA,
                 lnB=lnB, out_lnB=out_lnB)


def _ffn_cache(ff) -> PackedCache:
    c = ff.__dict__.get("_tavsr_packed")
    if c is None:
        c = ff.__dict__["_tavsr_packed"] = PackedCache()
    return c


def linear(x, w, bias, cache: PackedCache, key: str, act: int = ops.ACT_NONE, out_dtype=None):
    """act(x W^T + b) in the current compute mode (tiled tcgen05 GEMM)."""
    if _DTYPE == "tf32x3":
        w3 = cache.get(key + ":x3", [w], lambda: ops.split_tf32(w.detach(), "w"))
        return ops.gemm_bias_act(ops.split_tf32(x, "x"), w3, bias, act=act)
    return ops.gemm_bias_act(x, cache.weight(key, w), bias, act=act, out_dtype=out_dtype)


def linear_rowln(x, w, bias, cache: PackedCache, key: str, **kw):
    """Row-complete (N = 256) projection with the fused residual / LayerNorm epilogue in the current
    compute mode (single-operand form)."""
    if _DTYPE == "tf32x3":
        w3 = cache.get(key + ":x3", [w], lambda: ops.split_tf32(w.detach(), "w"))
        return ops.gemm_rowln(ops.split_tf32(x, "x"), w3, bias, **kw)
    return ops.gemm_rowln(x, cache.weight(key, w), bias, **kw)


def qkv_weights(attn, cache: PackedCache, key: str):
    srcs = [attn.linear_q.weight, attn.linear_k.weight, attn.linear_v.weight,
            attn.linear_q.bias, attn.linear_k.bias, attn.linear_v.bias]
    return cache.get(key, srcs, lambda: (
        torch.cat([attn.linear_q.weight, attn.linear_k.weight, attn.linear_v.weight], 0).contiguous(),
        torch.cat([attn.linear_q.bias, attn.linear_k.bias, attn.linear_v.bias], 0).contiguous()))


# the fused QKV projection and channel_proj1 + GELU of a two-branch block as one grouped launch
GROUP_PROJ = os.environ.get("TAVSR_GROUP_PROJ", "0") != "0"   # measured neutral at C2 (2.431 vs 2.427 ms): opt-in


def branch_projections(xa, xm, attn, cgmlp, cache: PackedCache):
    """(qkv, g) = (xa Wqkv^T + b, gelu(xm Wc1^T + bc1)) - one grouped launch when the mode allows
    (tf32 / bf16, equal input widths), else None (the callers then project separately)."""
    if not GROUP_PROJ or _DTYPE == "tf32x3":
        return None
    lin = cgmlp.channel_proj1[0]
    wqkv, bqkv = qkv_weights(attn, cache, "qkv")
    if xa.shape != xm.shape or wqkv.shape[1] != lin.weight.shape[1]:
        return None
    return ops.gemm_group2(xa, cache.weight("qkv.w", wqkv), bqkv, xm, cache.weight("conv.p1", lin.weight),
                           lin.bias, out_dtype=act_dtype())


def attention_ctx(xa, attn, pos_proj, lens, B, T, cache: PackedCache, key: str, qkv=None):
    """Fused QKV projection + rel-pos attention; returns ctx (B*T, d) in operand storage."""
    if qkv is None:
        wqkv, bqkv = qkv_weights(attn, cache, key)
        qkv = linear(xa, wqkv, bqkv, cache, key + ".w", out_dtype=act_dtype())
    u = attn.pos_bias_u.reshape(-1)
    v = attn.pos_bias_v.reshape(-1)
    return ops.relpos_attn(qkv, pos_proj, u, v, lens, B, T, attn.h)


def pos_projection(attn, pos_emb: torch.Tensor, cache: Optional[PackedCache] = None):
    """linear_pos(pos_emb): (2T-1, d).  Batch independent."""
    cache = cache if cache is not None else PackedCache()
    pe = operand(pos_emb.reshape(-1, pos_emb.shape[-1]).contiguous().float())
    return linear(pe, attn.linear_pos.weight, None, cache, f"wpos1_{id(attn)}", out_dtype=act_dtype())


def cgmlp_gated(xm, cgmlp, B, T, cache: PackedCache, key: str, g=None):
    """channel_proj1 + GELU + CSGU; returns u (B*T, C/2) ready for channel_proj2 (`g`: the
    GELU'd projection when the grouped launch already produced it)."""
    if cgmlp.csgu.linear is not None or cgmlp.csgu.gate_activation != "identity":
        raise NotImplementedError("use_linear_after_conv / non-identity gate_activation are not "
                                  "built on the B200 path (no shipped config uses them)")
    lin = cgmlp.channel_proj1[0]
    conv = cgmlp.csgu.conv
    cw = cache.get(key, [conv.weight], lambda: conv.weight.reshape(conv.weight.shape[0], -1).contiguous())
    if g is None:
        g = linear(xm, lin.weight, lin.bias, cache, key + ".p1", act=ops.ACT_GELU, out_dtype=act_dtype())
    return ops.csgu(g, cgmlp.csgu.norm.weight, cgmlp.csgu.norm.bias, cw, conv.bias, B, T,
                    eps=cgmlp.csgu.norm.eps)
