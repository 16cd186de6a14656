"""Kernel sequencing for one Branchformer block on the B200 path.

The layer modules own the parameters; this file turns them into launch sequences of the C-ABI
kernels (ops.py).  All activations are 2-D row-major (B*T, C) fp32 tensors ("tf32 mode": operands
are rounded to TF32 by TMA on load, accumulation and every non-GEMM step stay fp32).

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

    def get(self, key: str, sources, build):
        sig = tuple((s.data_ptr(), self._version(s), s.device) for s in sources)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = build()
        self._store[key] = (sig, val)
        return val


def require_inference(module: torch.nn.Module, *tensors) -> None:
    """The backward kernels of this path are not built yet: refuse loudly instead of silently
    detaching the graph."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise NotImplementedError(
            "tailored_avsr_b200: the CUDA backward kernels for the encoder are not built in this "
            "round; call the encoder under torch.no_grad() / torch.inference_mode() "
            "(see DESIGN.md, 'Training path').")
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("tailored_avsr_b200 runs on CUDA tensors only (there is no CPU "
                               "fallback); move the module and its inputs to a B200 device")


def lens_from_mask(mask: Optional[torch.Tensor], B: int, T: int, device) -> torch.Tensor:
    """(B,1,T) bool key mask -> int32 lengths.  The kernels implement prefix masks (what
    make_pad_mask produces, encoder.py:345)."""
    if mask is None:
        return torch.full((B,), T, dtype=torch.int32, device=device)
    return mask.reshape(B, -1).sum(dim=1).to(torch.int32)


def ffn_block(x, xn, ff, *, out_main, ln0=None, lnA=None, out_lnA=None, round_lnA=False,
              lnB=None, out_lnB=None, eps0=1e-12, alpha=0.5):
    """x + alpha * (W2 act(W1 xn + b1) + b2) with the trailing LayerNorms fused (x may be None: no
    residual).  One kernel with the hidden activation kept on chip when the shapes are the built
    ones (256 -> 2048 -> 256), otherwise (or with TAVSR_FFN_FUSED=0) the two-GEMM sequence."""
    if FFN_FUSED and tuple(ff.w_1.weight.shape) == (2048, 256):
        ops.ffn_fused(xn, ff.w_1.weight, ff.w_1.bias, ff.w_2.weight, ff.w_2.bias,
                      act_code(ff.activation_type), residual=x, alpha=alpha, ln0=ln0, eps0=eps0,
                      out_main=out_main, lnA=lnA, out_lnA=out_lnA, round_lnA=round_lnA,
                      lnB=lnB, out_lnB=out_lnB)
        return
    h = ops.gemm_bias_act(xn, ff.w_1.weight, ff.w_1.bias, act=act_code(ff.activation_type))
    ops.gemm_rowln(h, ff.w_2.weight, ff.w_2.bias, residual=x, alpha=alpha, ln0=ln0, eps0=eps0,
                   out_main=out_main, lnA=lnA, out_lnA=out_lnA, round_lnA=round_lnA,
                   lnB=lnB, out_lnB=out_lnB)


def qkv_weights(attn, cache: PackedCache, key: str):
    srcs = [attn.linear_q.weight, attn.linear_k.weight, attn.linear_v.weight,
            attn.linear_q.bias, attn.linear_k.bias, attn.linear_v.bias]
    return cache.get(key, srcs, lambda: (
        torch.cat([attn.linear_q.weight, attn.linear_k.weight, attn.linear_v.weight], 0).contiguous(),
        torch.cat([attn.linear_q.bias, attn.linear_k.bias, attn.linear_v.bias], 0).contiguous()))


def attention_ctx(xa, attn, pos_proj, lens, B, T, cache: PackedCache, key: str, dots=None):
    """Fused QKV projection + rel-pos attention; returns ctx (B*T, d), or (ctx, partial row dots
    (B*T, 2h, 2)) when `dots=(va, vb)` asks the attention epilogue for the learned_ave scores."""
    wqkv, bqkv = qkv_weights(attn, cache, key)
    qkv = ops.gemm_bias_act(xa, wqkv, bqkv)
    u = attn.pos_bias_u.reshape(-1)
    v = attn.pos_bias_v.reshape(-1)
    return ops.relpos_attn(qkv, pos_proj, u, v, lens, B, T, attn.h, dots=dots)


def pos_projection(attn, pos_emb: torch.Tensor):
    """linear_pos(pos_emb): (2T-1, d).  Batch independent."""
    return ops.gemm_bias_act(pos_emb.reshape(-1, pos_emb.shape[-1]), attn.linear_pos.weight, None)


# LayerNorm statistics of the CSGU gate half from the channel_proj1 GEMM epilogue (no stand-alone
# statistics kernel) and the learned_ave row dots from the attention / CSGU epilogues (no row_dots
# kernel)
# MEASURED (C2, CUDA-graph replay): both fusions LOSE inside the PDL-chained graph - the two small
# kernels they remove mostly overlap their neighbours there, while the extra epilogue work and the
# partial-summing merge-weights kernel sit on the critical path (3.86 ms -> 3.99-4.03 ms per step) -
# so they are opt-in; the ncu launch list (cold, serialised) shows the opposite, -18 us per block.
FUSE_STATS = os.environ.get("TAVSR_FUSE_STATS", "0") != "0"
FUSE_DOTS = os.environ.get("TAVSR_FUSE_DOTS", "0") != "0"


def cgmlp_gated(xm, cgmlp, B, T, cache: PackedCache, key: str, dots=None):
    """channel_proj1 + GELU + CSGU; returns u (B*T, C/2) ready for channel_proj2, or
    (u, partial row dots (B*T, C/256, 2)) when `dots=(va, vb)` is given."""
    if cgmlp.csgu.linear is not None or cgmlp.csgu.gate_activation != "identity":
        raise NotImplementedError("use_linear_after_conv / non-identity gate_activation are not "
                                  "built on the B200 path (no shipped config uses them)")
    lin = cgmlp.channel_proj1[0]
    conv = cgmlp.csgu.conv
    cw = cache.get(key, [conv.weight], lambda: conv.weight.reshape(conv.weight.shape[0], -1).contiguous())
    Ch = lin.weight.shape[0] // 2
    if FUSE_STATS and Ch % 128 == 0 and Ch // 64 <= 16:
        g, st, n_part, pw = ops.gemm_bias_act_stats(xm, lin.weight, lin.bias, ops.ACT_GELU, Ch)
        u, d = ops.csgu_fused(g, cgmlp.csgu.norm.weight, cgmlp.csgu.norm.bias, cw, conv.bias, B, T,
                              st, n_part, pw, eps=cgmlp.csgu.norm.eps, dots=dots)
        return u if dots is None else (u, d)
    if dots is not None and Ch % 128 == 0:
        g = ops.gemm_bias_act(xm, lin.weight, lin.bias, act=ops.ACT_GELU)
        return ops.csgu_fused(g, cgmlp.csgu.norm.weight, cgmlp.csgu.norm.bias, cw, conv.bias, B, T,
                              None, 0, 0, eps=cgmlp.csgu.norm.eps, dots=dots)
    g = ops.gemm_bias_act(xm, lin.weight, lin.bias, act=ops.ACT_GELU)
    u = ops.csgu(g, cgmlp.csgu.norm.weight, cgmlp.csgu.norm.bias, cw, conv.bias, B, T,
                 eps=cgmlp.csgu.norm.eps)
    if dots is None:
        return u
    d, _ = ops.row_dots(u, dots[0], dots[1])
    return u, d.view(-1, 1, 2)
