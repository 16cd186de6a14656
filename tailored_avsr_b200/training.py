"""Training path of the Branchformer encoder on the B200 kernels: forward that keeps what the
backward needs, hand-composed backward, wired into autograd as ONE node per encoder block.

Reference being differentiated: `MyBranchformerEncoderLayer.forward`
(src/encoder/branchformer/encoder_layer.py:153-321) inside `MyBranchformerEncoder.forward`
(src/encoder/branchformer/encoder.py:324-412), trained every step by avsr_main.py:27-57.

Why one autograd node per block: autograd runs the blocks' backward nodes last-to-first and hands
every parameter gradient to its AccumulateGrad node as soon as the block's node returns, so the
bucketed gradient all-reduce (parallel.GradBucketReducer, hooked per parameter) overlaps the
backward of the earlier blocks - the DDP pattern SURVEY.md §8e asks for.

Forward = the un-fused kernel sequence (the fused FFN keeps its 2048-wide hidden on chip and the
folded merge never forms x1 / x2: both would have to be recomputed for the weight gradients):
    xn0 = LN_ffmac(x)                      z1 = xn0 W1^T + b1          h1 = swish(z1)
    x_a = x + .5 (h1 W2^T + b2)            xa = LN_mha(x_a), xm = LN_mlp(x_a)
    qkv = xa Wqkv^T + b                    ctx, lse = relpos_attn(qkv, linear_pos(pos_emb), u, v)
    x1 = ctx Wo^T + bo                     zc = xm Wc1^T + bc1, hc = gelu(zc), u = csgu(hc), x2 = u Wc2^T + b
    (w1, w2) = learned_ave(x1, x2)         x_b = x_a + c (w1 x1 + w2 x2) Wm^T + bm,  xf = LN_ff(x_b)
    z2, h2 likewise                        y0 = x_b + .5 (h2 W2^T + b2),   y = LN_final(y0)
Kept per block: x, xn0, z1, x_a, xa, xm, qkv, pos_proj, ctx, lse, x1, zc, u, CSGU (mean, rstd), x2,
w1, w2, x_b, xf, z2, y0 (~0.6 GB at the C2 shape).  Backward = oracle/manual_backward.py's
composition (equal to autograd on the reference's own modules: tests/golden/grad_*.npz) on the
kernels of ops_backward.py: dgrad / wgrad on the tcgen05 GEMM, LayerNorm / activation / CSGU /
merge / rel-pos attention backward kernels.  Storage is fp32 and the products run on TF32 operands
in every compute mode (the bf16 mode is an inference mode).

Training-only randomness follows the reference's RNG streams: stochastic depth draws
`torch.rand(1).item()` at layer entry (encoder_layer.py:180-182), attention-branch drop draws it
inside the learned_ave merge (:233-239), in that order, layer by layer (host RNG).  Dropout: every
site of the reference is built - per block, in the reference's call order: FFN hidden and output
(macaron), attention probabilities, x1, CSGU output, x2, merge output, FFN hidden and output
(encoder_layer.py:194, 208-224, 228-306, 314 and the espnet leaves behind them); in the `linear`
input layer the Dropout after its LayerNorm and the two of RelPositionalEncoding (x and pos_emb).
Keep-masks are drawn per call from `torch.nn.functional.dropout` on a ones tensor of the
reference's shape on the device - the same generator stream, kernel and shapes the reference's
nn.Dropout modules consume - in that order, then applied by our kernels (elementwise keep * 1/(1-p)
masks; the attention kernels take a byte keep-mask).  `set_dropout_source` swaps the generator for a
deterministic one (tests: the same masks are injected into the live reference).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch

from . import engine, ops
from . import ops_backward as ob

F32 = torch.float32


def wants_grad(module: torch.nn.Module, *tensors) -> bool:
    """True when this call has to build an autograd graph (grad mode on and a parameter or an input
    asks for gradients)."""
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


_DROP_SOURCE = None


def set_dropout_source(fn) -> None:
    """fn(shape, p, device) -> fp32 mask of `shape` holding 0 or 1 / (1 - p); None restores the
    default (torch's own dropout kernel on the device generator)."""
    global _DROP_SOURCE
    _DROP_SOURCE = fn


def draw_mask(shape, p: float, device) -> Optional[torch.Tensor]:
    """One dropout mask (already scaled by 1 / (1 - p)), or None when the site is inactive."""
    if not p > 0:
        return None
    if p >= 1:
        raise NotImplementedError("dropout with p >= 1 is not built")
    if _DROP_SOURCE is not None:
        return _DROP_SOURCE(tuple(shape), float(p), device).to(device=device, dtype=F32).contiguous()
    return torch.nn.functional.dropout(torch.ones(tuple(shape), device=device, dtype=F32), float(p), True)


def _mul(x: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """x * mask on our elementwise kernel (identity activation of tavsr_act_fwd)."""
    return x if mask is None else ob.act_fwd(x, ops.ACT_NONE, mask=mask.view(x.shape))


def _keep_bytes(mask: torch.Tensor, T: int) -> torch.Tensor:
    """(B, H, T, T) scaled fp32 mask -> (B, H, T, Tp) uint8 keep flags, Tp = T rounded up to 128
    (the layout tavsr_relpos_attn_fwd_dropout reads; host-side plumbing, once per layer call)."""
    Tp = (T + 127) // 128 * 128
    keep = torch.zeros(mask.shape[:3] + (Tp,), device=mask.device, dtype=torch.uint8)
    keep[..., :T] = mask != 0
    return keep


def _named(layer: torch.nn.Module) -> Tuple[List[str], List[torch.nn.Parameter]]:
    names, params = [], []
    for n, p in layer.named_parameters():
        names.append(n)
        params.append(p)
    return names, params


def _acc(g: Dict[str, torch.Tensor], key: str, val: torch.Tensor) -> None:
    g[key] = g[key] + val if key in g else val


def _lin_bwd(g, x_in, lin_w, dy, wname: Optional[str], bname: Optional[str], need_dx=True, xT=None,
             dyT=None):
    dx, dw, db = ob.linear_bwd(x_in, lin_w, dy, need_dx=need_dx, need_dw=wname is not None, xT=xT,
                               dyT=dyT)
    if wname is not None:
        _acc(g, wname, dw)
        if bname is not None:
            _acc(g, bname, db)
    return dx


def _ln_bwd(g, x, norm, dy, prefix: str, eps: float = 1e-12, dres=None):
    dx, dg, db = ob.layernorm_bwd(x, norm.weight, dy, eps=eps, dres=dres)
    _acc(g, prefix + ".weight", dg)
    _acc(g, prefix + ".bias", db)
    return dx


def _half(t: torch.Tensor) -> torch.Tensor:
    """0.5 * t on our elementwise kernel (the ff_scale of the macaron FFNs)."""
    half, zero = _scalars(t.device)
    return ops.scale_add_rows(t, t, half, zero, max(1, t.shape[0]))


_SC = {}


def _scalars(device):
    s = _SC.get(device)
    if s is None:
        s = _SC[device] = (torch.full((1,), 0.5, device=device), torch.zeros(1, device=device))
    return s


# ---------------------------------------------------------------------------------------------------
# position-wise feed-forward (espnet PositionwiseFeedForward + the residual line around it)
# ---------------------------------------------------------------------------------------------------
def _residual_ln(residual, t, alpha: float, out_main, lnA=None, out_lnA=None, lnB=None, out_lnB=None):
    """out_main = residual + alpha * t, then up to two LayerNorms of it: the un-fused form of the
    row-complete GEMM epilogue, used where a dropout mask sits between the GEMM and the residual."""
    dev = t.device
    one = ops._cast_scalars(dev)[0]
    a = _scalars(dev)[0] if alpha == 0.5 else torch.full((1,), float(alpha), device=dev)
    ops.scale_add_rows(residual, t, one, a, max(1, t.shape[0]), out=out_main)
    if lnA is not None:
        if lnB is not None:
            ops.layernorm(out_main, lnA[0], lnA[1], eps=1e-12, out=out_lnA, gB=lnB[0], bB=lnB[1],
                          outB=out_lnB)
        else:
            ops.layernorm(out_main, lnA[0], lnA[1], eps=1e-12, out=out_lnA)


def _ffn_forward(xn, ff, residual, out_main, lnA=None, out_lnA=None, lnB=None, out_lnB=None,
                 mask_h=None, mask_o=None):
    act = engine.act_code(ff.activation_type)
    z = ops.gemm_bias_act(xn, ff.w_1.weight, ff.w_1.bias)
    h = ob.act_fwd(z, act, mask=mask_h)
    if mask_o is None:
        ops.gemm_rowln(h, ff.w_2.weight, ff.w_2.bias, residual=residual, alpha=0.5, out_main=out_main,
                       lnA=lnA, out_lnA=out_lnA, lnB=lnB, out_lnB=out_lnB)
    else:
        t = _mul(ops.gemm_bias_act(h, ff.w_2.weight, ff.w_2.bias), mask_o)
        _residual_ln(residual, t, 0.5, out_main, lnA, out_lnA, lnB, out_lnB)
    return z


def _ffn_backward(g, xn, z, ff, dout, prefix: str, mask_h=None, mask_o=None):
    """dout = d loss / d (residual + 0.5 drop_o(ffn(xn))); returns d loss / d xn."""
    act = engine.act_code(ff.activation_type)
    dhalf = _mul(_half(dout), mask_o)
    # h is recomputed from z (only z is kept) and is needed only as the wgrad operand h^T; dz is
    # needed row-major (dgrad) and transposed (wgrad): both come out of fused one-pass kernels
    hT = ob.act_fwd_t(z, act, mask=mask_h)
    dh = _lin_bwd(g, None, ff.w_2.weight, dhalf, prefix + ".w_2.weight", prefix + ".w_2.bias", xT=hT)
    dz, dzT = ob.act_bwd_t(z, _mul(dh, mask_h), act)
    return _lin_bwd(g, xn, ff.w_1.weight, dz, prefix + ".w_1.weight", prefix + ".w_1.bias", dyT=dzT)


# ---------------------------------------------------------------------------------------------------
# one Branchformer block
# ---------------------------------------------------------------------------------------------------
def block_forward(L, aux, x: torch.Tensor) -> Tuple[torch.Tensor, dict]:
    B, T = aux.B, aux.T
    M, d = B * T, L.size
    dev = x.device
    new = lambda c=d: torch.empty((M, c), device=dev, dtype=F32)  # noqa: E731
    sv = {"x": x}
    drp = aux.drop
    two = L.use_two_branches
    learned = two and L.merge_method == "learned_ave" and not aux.drop_attn
    # ---- macaron FFN ----
    sv["xn0"] = ops.layernorm(x, L.norm_ff_macaron.weight, L.norm_ff_macaron.bias, eps=1e-12)
    x_a = new()
    xa = new() if L.attn is not None else None
    xm = new() if L.cgmlp is not None else None
    lnA = (L.norm_mha.weight, L.norm_mha.bias) if L.attn is not None else None
    lnB = (L.norm_mlp.weight, L.norm_mlp.bias) if L.cgmlp is not None else None
    if lnA is None:
        sv["z1"] = _ffn_forward(sv["xn0"], L.feed_forward_macaron, x, x_a, lnA=lnB, out_lnA=xm,
                                mask_h=drp.get("ffm_h"), mask_o=drp.get("ffm_o"))
    else:
        sv["z1"] = _ffn_forward(sv["xn0"], L.feed_forward_macaron, x, x_a, lnA=lnA, out_lnA=xa,
                                lnB=lnB, out_lnB=xm, mask_h=drp.get("ffm_h"), mask_o=drp.get("ffm_o"))
    sv.update(x_a=x_a, xa=xa, xm=xm)
    x1 = x2 = d1 = d2 = None
    # ---- attention branch ----
    if L.attn is not None:
        A = L.attn
        if aux.pos2d is None:
            raise NotImplementedError("attention without relative positional embedding is not built")
        wqkv = torch.cat([A.linear_q.weight, A.linear_k.weight, A.linear_v.weight], 0)
        bqkv = torch.cat([A.linear_q.bias, A.linear_k.bias, A.linear_v.bias], 0)
        qkv = ops.gemm_bias_act(xa, wqkv, bqkv)
        pp = ops.gemm_bias_act(aux.pos2d, A.linear_pos.weight, None)
        lse = torch.empty((B, A.h, T), device=dev, dtype=F32)
        ctx = ops.relpos_attn(qkv, pp, A.pos_bias_u.reshape(-1), A.pos_bias_v.reshape(-1), aux.lens,
                              B, T, A.h, round_out=False, lse=lse, drop=drp.get("att"))
        x1 = new()
        dots = None
        if learned:
            d1 = torch.empty((M, 2), device=dev, dtype=F32)
            dots = (L.pooling_proj1.weight.reshape(-1), L.weight_proj1.weight.reshape(-1))
        if drp.get("x1") is None:
            ops.gemm_rowln(ctx, A.linear_out.weight, A.linear_out.bias, out_main=x1, dots=dots, dots_out=d1)
        else:   # x1 = dropout(linear_out(ctx)) (:212); the pooling scores are taken on the dropped x1
            ops.gemm_rowln(ctx, A.linear_out.weight, A.linear_out.bias, out_main=x1)
            x1 = _mul(x1, drp["x1"])
            if learned:
                d1 = ops.row_dots(x1, dots[0], dots[1])[0]
        sv.update(wqkv=wqkv, qkv=qkv, pp=pp, lse=lse, ctx=ctx, x1=x1)
    # ---- cgMLP branch ----
    if L.cgmlp is not None:
        Cg = L.cgmlp
        if Cg.csgu.linear is not None or Cg.csgu.gate_activation != "identity":
            raise NotImplementedError("use_linear_after_conv / non-identity gate are not built")
        lin = Cg.channel_proj1[0]
        conv = Cg.csgu.conv
        zc = ops.gemm_bias_act(xm, lin.weight, lin.bias)
        hc = ob.act_fwd(zc, ops.ACT_GELU)
        stats = torch.empty((M, 2), device=dev, dtype=F32)
        u = ops.csgu(hc, Cg.csgu.norm.weight, Cg.csgu.norm.bias, conv.weight.reshape(conv.weight.shape[0], -1),
                     conv.bias, B, T, eps=Cg.csgu.norm.eps, round_out=False, stats=stats)
        u = _mul(u, drp.get("csgu"))                      # espnet cgmlp.py: dropout(x_r * x_g)
        x2 = new()
        dots = None
        if learned:
            d2 = torch.empty((M, 2), device=dev, dtype=F32)
            dots = (L.pooling_proj2.weight.reshape(-1), L.weight_proj2.weight.reshape(-1))
        if drp.get("x2") is None:
            ops.gemm_rowln(u, Cg.channel_proj2.weight, Cg.channel_proj2.bias, out_main=x2, dots=dots,
                           dots_out=d2)
        else:
            ops.gemm_rowln(u, Cg.channel_proj2.weight, Cg.channel_proj2.bias, out_main=x2)
            x2 = _mul(x2, drp["x2"])
            if learned:
                d2 = ops.row_dots(x2, dots[0], dots[1])[0]
        sv.update(zc=zc, stats=stats, u=u, x2=x2)
    # ---- merge (:227-309) ----
    x_b, xf = new(), new()
    lnF = (L.norm_ff.weight, L.norm_ff.bias)
    mp = L.merge_proj
    ident = isinstance(mp, torch.nn.Identity)
    if ident:
        # single branch added straight onto the residual: x_b = x_a + c * branch (the tailored AV
        # layer, tailored/encoder_layer.py:185-207; its branch dropout is the x1 / x2 site)
        if two:
            raise NotImplementedError("two-branch block with merge_proj=Identity is not built")
        _residual_ln(x_a, x2 if L.attn is None else x1, aux.stoch, x_b, lnF, xf)
    elif two and L.merge_method in ("learned_ave", "fixed_ave"):
        if L.merge_method == "learned_ave":
            if learned:
                scal = torch.cat([L.pooling_proj1.bias, L.pooling_proj2.bias, L.weight_proj1.bias,
                                  L.weight_proj2.bias]).float()
                w1, w2 = ops.merge_weights_dev(d1, d2, aux.lens, scal, d, B, T)
            else:   # attention branch dropped for this step (:233-239)
                w1 = torch.zeros((B,), device=dev, dtype=F32)
                w2 = torch.ones((B,), device=dev, dtype=F32)
            L.weight_global, L.weight_local = w1.detach().view(B, 1, 1), w2.detach().view(B, 1, 1)
        else:
            w1 = torch.full((B,), 1.0 - L.cgmlp_weight, device=dev, dtype=F32)
            w2 = torch.full((B,), float(L.cgmlp_weight), device=dev, dtype=F32)
        if drp.get("merge") is None:
            ops.gemm_rowln(x1, mp.weight, mp.bias, x2=x2, rowscale=(w1, w2), rows_per_seg=T,
                           residual=x_a, alpha=aux.stoch, out_main=x_b, lnA=lnF, out_lnA=xf)
        else:
            msrc = ops.scale_add_rows(x1, x2, w1, w2, T)
        sv.update(w1=w1, w2=w2)
    elif two:   # concat
        msrc = torch.cat([x1, x2], 1)
        if drp.get("merge") is None:
            ops.gemm_rowln(msrc, mp.weight, mp.bias, residual=x_a, alpha=aux.stoch, out_main=x_b,
                           lnA=lnF, out_lnA=xf)
    else:
        msrc = x2 if L.attn is None else x1
        if drp.get("merge") is None:
            ops.gemm_rowln(msrc, mp.weight, mp.bias, residual=x_a, alpha=aux.stoch, out_main=x_b,
                           lnA=lnF, out_lnA=xf)
    if drp.get("merge") is not None and not ident:   # x_b = x_a + c * dropout(merge_proj(.)) (:228-306)
        t = _mul(ops.gemm_bias_act(msrc, mp.weight, mp.bias), drp["merge"])
        _residual_ln(x_a, t, aux.stoch, x_b, lnF, xf)
    sv.update(x_b=x_b, xf=xf, learned=learned)
    # ---- FFN + norm_final ----
    y0, y = new(), new()
    sv["z2"] = _ffn_forward(xf, L.feed_forward, x_b, y0, lnA=(L.norm_final.weight, L.norm_final.bias),
                            out_lnA=y, mask_h=drp.get("ff_h"), mask_o=drp.get("ff_o"))
    sv["y0"] = y0
    return y, sv


def block_backward(L, aux, sv: dict, dy: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    B, T = aux.B, aux.T
    M, d = B * T, L.size
    dev = dy.device
    g: Dict[str, torch.Tensor] = {}
    drp = aux.drop
    two = L.use_two_branches
    # ---- norm_final, FFN, norm_ff ----
    dy0 = _ln_bwd(g, sv["y0"], L.norm_final, dy, "norm_final")
    dxf = _ffn_backward(g, sv["xf"], sv["z2"], L.feed_forward, dy0, "feed_forward",
                        mask_h=drp.get("ff_h"), mask_o=drp.get("ff_o"))
    dx_b = _ln_bwd(g, sv["x_b"], L.norm_ff, dxf, "norm_ff", dres=dy0)          # residual joins here
    # ---- merge ----
    mp = L.merge_proj
    dmo = dx_b
    if aux.stoch != 1.0:   # x_b = x_a + c * (merge_proj(m)): the projection sees c * dx_b
        c = torch.full((1,), float(aux.stoch), device=dev)
        dmo = ops.scale_add_rows(dx_b, dx_b, c, _scalars(dev)[1], M)
    ident = isinstance(mp, torch.nn.Identity)
    if not ident:
        dmo = _mul(dmo, drp.get("merge"))
    x1, x2 = sv.get("x1"), sv.get("x2")
    dx1 = dx2 = None
    if ident:
        dx1, dx2 = (None, dmo) if L.attn is None else (dmo, None)
    elif two and L.merge_method in ("learned_ave", "fixed_ave"):
        w1, w2 = sv["w1"], sv["w2"]
        m = ops.scale_add_rows(x1, x2, w1, w2, T)
        dm = _lin_bwd(g, m, mp.weight, dmo, "merge_proj.weight", "merge_proj.bias")
        if sv["learned"]:
            scal = torch.cat([L.pooling_proj1.bias, L.weight_proj1.bias, L.pooling_proj2.bias,
                              L.weight_proj2.bias]).float()
            dx1, dx2, gr = ob.merge_learned_ave_bwd(
                x1, x2, dm, aux.lens, L.pooling_proj1.weight.reshape(-1), L.weight_proj1.weight.reshape(-1),
                L.pooling_proj2.weight.reshape(-1), L.weight_proj2.weight.reshape(-1), scal, B, T)
            D = d
            for k, name in enumerate(("pooling_proj1", "weight_proj1", "pooling_proj2", "weight_proj2")):
                _acc(g, name + ".weight", gr[k * D:(k + 1) * D].reshape(1, D))
                _acc(g, name + ".bias", gr[4 * D + k:4 * D + k + 1])
        else:
            zero = torch.zeros((B,), device=dev, dtype=F32)
            dx1 = ops.scale_add_rows(dm, dm, w1, zero, T)
            dx2 = ops.scale_add_rows(dm, dm, w2, zero, T)
    elif two:   # concat
        cat = torch.cat([x1, x2], 1)
        dcat = _lin_bwd(g, cat, mp.weight, dmo, "merge_proj.weight", "merge_proj.bias")
        dx1, dx2 = dcat[:, :d], dcat[:, d:]
    else:
        xs = x2 if L.attn is None else x1
        dxs = _lin_bwd(g, xs, mp.weight, dmo, "merge_proj.weight", "merge_proj.bias")
        dx1, dx2 = (None, dxs) if L.attn is None else (dxs, None)
    dx_a = dx_b
    # ---- attention branch ----
    if L.attn is not None:
        A = L.attn
        dx1 = _mul(dx1, drp.get("x1"))
        dctx = _lin_bwd(g, sv["ctx"], A.linear_out.weight, dx1, "attn.linear_out.weight", "attn.linear_out.bias")
        dqkv, dpos, du, dv = ob.relpos_attn_bwd(sv["qkv"], sv["pp"], A.pos_bias_u.reshape(-1),
                                                A.pos_bias_v.reshape(-1), aux.lens, sv["ctx"], dctx,
                                                sv["lse"], B, T, A.h, drop=drp.get("att"))
        _acc(g, "attn.pos_bias_u", du.view(A.h, A.d_k))
        _acc(g, "attn.pos_bias_v", dv.view(A.h, A.d_k))
        # d linear_pos.weight = dpos^T . pos_emb (reduction over the 2T-1 relative positions)
        _acc(g, "attn.linear_pos.weight",
             ops.gemm_bias_act(ob.transpose_2d(dpos, pad=True), aux.pos2d_T, None))
        dxa, dwqkv, dbqkv = ob.linear_bwd(sv["xa"], sv["wqkv"], dqkv)
        D = d
        for k, name in enumerate(("q", "k", "v")):
            _acc(g, f"attn.linear_{name}.weight", dwqkv[k * D:(k + 1) * D])
            _acc(g, f"attn.linear_{name}.bias", dbqkv[k * D:(k + 1) * D])
        dx_a = _ln_bwd(g, sv["x_a"], L.norm_mha, dxa, "norm_mha", dres=dx_a)
    # ---- cgMLP branch ----
    if L.cgmlp is not None:
        Cg = L.cgmlp
        lin, conv = Cg.channel_proj1[0], Cg.csgu.conv
        dx2 = _mul(dx2, drp.get("x2"))
        du_ = _mul(_lin_bwd(g, sv["u"], Cg.channel_proj2.weight, dx2, "cgmlp.channel_proj2.weight",
                            "cgmlp.channel_proj2.bias"), drp.get("csgu"))
        hc = ob.act_fwd(sv["zc"], ops.ACT_GELU)
        dhc, dng, dnb, dcw, dcb = ob.csgu_bwd(hc, Cg.csgu.norm.weight, Cg.csgu.norm.bias,
                                              conv.weight.reshape(conv.weight.shape[0], -1), conv.bias,
                                              sv["stats"], du_, B, T, eps=Cg.csgu.norm.eps)
        _acc(g, "cgmlp.csgu.norm.weight", dng)
        _acc(g, "cgmlp.csgu.norm.bias", dnb)
        _acc(g, "cgmlp.csgu.conv.weight", dcw.view_as(conv.weight))
        _acc(g, "cgmlp.csgu.conv.bias", dcb)
        dzc, dzcT = ob.act_bwd_t(sv["zc"], dhc, ops.ACT_GELU)
        dxm = _lin_bwd(g, sv["xm"], lin.weight, dzc, "cgmlp.channel_proj1.0.weight",
                       "cgmlp.channel_proj1.0.bias", dyT=dzcT)
        dx_a = _ln_bwd(g, sv["x_a"], L.norm_mlp, dxm, "norm_mlp", dres=dx_a)
    # ---- macaron FFN ----
    dxn0 = _ffn_backward(g, sv["xn0"], sv["z1"], L.feed_forward_macaron, dx_a, "feed_forward_macaron",
                         mask_h=drp.get("ffm_h"), mask_o=drp.get("ffm_o"))
    dx = _ln_bwd(g, sv["x"], L.norm_ff_macaron, dxn0, "norm_ff_macaron", dres=dx_a)
    return dx, g


class _BlockFn(torch.autograd.Function):
    """One Branchformer block as one autograd node (forward: block_forward, backward: block_backward)."""

    @staticmethod
    def forward(ctx, L, aux, names, x, *params):
        y, sv = block_forward(L, aux, x.contiguous())
        ctx.L, ctx.aux, ctx.sv, ctx.names = L, aux, sv, names
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, g = block_backward(ctx.L, ctx.aux, ctx.sv, dy.contiguous())
        ctx.sv = None   # free the saved activations as soon as the block is done
        grads = []
        for n, p in zip(ctx.names, ctx.params):
            gp = g.get(n)
            grads.append(gp.reshape(p.shape) if gp is not None and p.requires_grad else None)
        return (None, None, None, dx) + tuple(grads)


class _LayerNormFn(torch.autograd.Function):
    """y = scale * LayerNorm(x) (after_norm; the `linear` input layer's LayerNorm x sqrt(d))."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, scale):
        ctx.save_for_backward(x, weight)
        ctx.eps, ctx.scale = eps, scale
        return ops.layernorm(x, weight, bias, eps=eps, scale=scale)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        s = ctx.scale
        gam = weight if s == 1.0 else weight * s        # d/dx of s * LN(x) == LN-backward with s * gamma
        dx, dg, db = ob.layernorm_bwd(x, gam.contiguous(), dy.contiguous(), eps=ctx.eps)
        if s != 1.0:
            dg, db = dg * s, db * s
        return dx, dg, db, None, None


class _MaskFn(torch.autograd.Function):
    """y = x * mask (a dropout site outside the blocks)."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return _mul(x.contiguous(), mask)

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        return _mul(dy.contiguous(), mask), None


class _CondFn(torch.autograd.Function):
    """InterCTC self-conditioning, x + conditioning_layer(softmax(ctc_lo(tap))) (encoder.py:393-401;
    ctc.py:160-168), as one autograd node: the CTC head and vocab-residual kernels forward; backward
    = conditioning_layer's dgrad / wgrad over the vocabulary on the CTC head kernels, the softmax
    backward, then the CTC head backward (gradients for x, the tap, ctc_lo and conditioning_layer)."""

    @staticmethod
    def forward(ctx, x, tap, wc, bc, wl, bl):
        if wc.shape[0] > 64:
            raise NotImplementedError("InterCTC conditioning on the training path is built for odim <= 64")
        _, prob, _ = ops.ctc_head(tap.contiguous(), wc, bc, want_logp=False, want_prob=True)
        out, _ = ops.vocab_residual(x.contiguous(), prob, wl.contiguous(), bl)
        ctx.save_for_backward(tap, prob, wc, wl)
        return out

    @staticmethod
    def backward(ctx, dout):
        tap, prob, wc, wl = ctx.saved_tensors
        dout = dout.contiguous()
        V = wc.shape[0]
        wlT = ob.transpose_2d(wl.contiguous())                        # (V, d)
        zero_b = torch.zeros((V,), device=dout.device, dtype=F32)
        dp = ops.ctc_head(dout, wlT, zero_b, want_logp=False, want_logits=True)[3]     # dout . Wl
        _, dwlT, _ = ops.ctc_head_bwd(prob, dout, wlT)                # (V, d) = prob^T . dout
        dl = ob.softmax_bwd(prob, dp)
        dtap, dwc, dbc = ops.ctc_head_bwd(dl, tap.contiguous(), wc.contiguous())
        return dout, dtap, dwc, dbc, ob.transpose_2d(dwlT), ob.col_sums(dout)


class _ScaleFn(torch.autograd.Function):
    """y = s * x on our elementwise kernel (the positional encoding's x * sqrt(d))."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.s = float(s)
        sc = torch.full((1,), ctx.s, device=x.device)
        return ops.scale_add_rows(x, x, sc, _scalars(x.device)[1], max(1, x.shape[0]))

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        sc = torch.full((1,), ctx.s, device=dy.device)
        return ops.scale_add_rows(dy, dy, sc, _scalars(dy.device)[1], max(1, dy.shape[0])), None


class _Conv2dFrontFn(torch.autograd.Function):
    """espnet Conv2dSubsampling up to its Linear (encoder.py:149-155): x (B, Tin, F) ->
    Linear(flatten(relu(conv2(relu(conv1(x)))))) as (B*T2, d).  Forward = the inference kernels (conv1
    on the fly inside the im2col writer, conv2 and the projection on the tcgen05 GEMM); backward =
    dgrad / wgrad of the two GEMMs + the col2im gather with conv1's ReLU mask and weight reduction
    (tavsr_conv2d_sub_bwd).  The input features get no gradient."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, wl, bl):
        B, Tin, Fin = x.shape
        C = w1.shape[0]
        T2, F2 = ((Tin - 1) // 2 - 1) // 2, ((Fin - 1) // 2 - 1) // 2
        w1p = w1.reshape(C, 9).contiguous().float()
        w2p = w2.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous().float()
        wlp = wl.view(-1, C, F2).permute(0, 2, 1).reshape(-1, F2 * C).contiguous().float()
        xc = x.contiguous().float()
        A = ops.conv2d_sub_im2col(xc, w1p, b1)
        h2 = ops.gemm_bias_act(A, w2p, b2, act=ops.ACT_RELU)          # relu'(z) == (h > 0)
        y = ops.gemm_bias_act(h2.view(B * T2, F2 * C), wlp, bl)
        ctx.save_for_backward(xc, w1p, b1, A, h2, w2p, wlp)
        ctx.dims = (B, T2, F2, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w1p, b1, A, h2, w2p, wlp = ctx.saved_tensors
        B, T2, F2, C = ctx.dims
        d = wlp.shape[0]
        dh2, dwlp, dbl = ob.linear_bwd(h2.view(B * T2, F2 * C), wlp, dy.contiguous())
        dz2 = ob.act_bwd(h2, dh2.view(B * T2 * F2, C), ops.ACT_RELU)
        dA, dw2p, db2 = ob.linear_bwd(A, w2p, dz2)
        dw1, db1 = ob.conv2d_sub_bwd(xc, w1p, b1, dA)
        dwl = dwlp.view(d, F2, C).permute(0, 2, 1).reshape(d, C * F2)
        dw2 = dw2p.view(C, 3, 3, C).permute(0, 3, 1, 2)
        return None, dw1.view(C, 1, 3, 3), db1, dw2, db2, dwl, dbl


def conv2d_front_forward(embed, x_in: torch.Tensor):
    """Conv2dSubsampling / Conv2dSubsamplingWOPosEnc parameters -> (B*T2, d) projection output."""
    conv = embed.conv
    lin = embed.out[0] if isinstance(embed.out, torch.nn.Sequential) else embed.out
    return _Conv2dFrontFn.apply(x_in, conv[0].weight, conv[0].bias, conv[2].weight, conv[2].bias,
                                lin.weight, lin.bias)


def posenc_dropouts(pe, x2d: torch.Tensor, pos_emb: torch.Tensor, B: int, T: int, active: bool):
    """espnet RelPositionalEncoding.forward in train mode: dropout(x * sqrt(d)) (the scaling is done
    by the caller), dropout(pos_emb), in that order."""
    d = x2d.shape[1]
    if active:
        m = draw_mask((B, T, d), float(pe.dropout_rate), x2d.device)
        if m is not None:
            x2d = _MaskFn.apply(x2d, m.view(B * T, d))
        m = draw_mask(tuple(pos_emb.shape), float(pe.dropout_rate), x2d.device)
        if m is not None:
            pos_emb = _mul(pos_emb.reshape(-1, d).contiguous().float(), m.view(-1, d)).view(pos_emb.shape)
    return x2d, pos_emb


def avsr_embed_forward(E, xs_pad: torch.Tensor, masks: torch.Tensor):
    """Training form of DefaultEmbeddingLayerForAVSR.apply_embed_layer (default.py:139-153):
    conv2d (Conv2dSubsamplingWOPosEnc), linear (Linear + LayerNorm + Dropout), None / Linear."""
    from .embedding_for_avsr.default import Conv2dSubsamplingWOPosEnc
    d = E._output_size
    if isinstance(E.embed, Conv2dSubsamplingWOPosEnc):
        B, Tin, _ = xs_pad.shape
        T = ((Tin - 1) // 2 - 1) // 2
        x = conv2d_front_forward(E.embed, xs_pad)
        return x.view(B, T, d), masks[:, :, :-2:2][:, :, :-2:2]
    if E.embed is None:
        return xs_pad, masks
    B, T, Fin = xs_pad.shape
    x2 = xs_pad.reshape(B * T, Fin).contiguous().float()
    if isinstance(E.embed, torch.nn.Sequential):
        lin, ln, dr = E.embed[0], E.embed[1], E.embed[2]
        x = _LayerNormFn.apply(_LinearFn.apply(x2, lin.weight, lin.bias), ln.weight, ln.bias, ln.eps, 1.0)
        if E.training:
            m = draw_mask((B, T, d), float(dr.p), x.device)
            if m is not None:
                x = _MaskFn.apply(x, m.view(B * T, d))
    else:
        x = _LinearFn.apply(x2, E.embed.weight, E.embed.bias)
    return x.view(B, T, d), masks


def avsr_posenc_forward(E, xs_pad: torch.Tensor):
    """Training form of apply_pos_enc (default.py:156-162): (dropout(x * sqrt(d)), dropout(pos_emb))."""
    B, T, d = xs_pad.shape
    x = _ScaleFn.apply(xs_pad.reshape(B * T, d).contiguous().float(), math.sqrt(d))
    pos_emb = E.pos_enc.pos_emb(T, xs_pad.device)
    x, pos_emb = posenc_dropouts(E.pos_enc, x, pos_emb, B, T, E.training)
    return x.view(B, T, d), pos_emb


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM (the `linear` input layer)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.need_dx = x.requires_grad
        return ops.gemm_bias_act(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx, dw, db = ob.linear_bwd(x, weight, dy.contiguous(), need_dx=ctx.need_dx)
        return dx, dw, db


# ---------------------------------------------------------------------------------------------------
# TailoredEncoderLayer (src/encoder/audiovisual/tailored/encoder_layer.py:118-260): per stream
#   x = x + .5 drop(FFN_mac(LN_mac(x)));  x = x + c drop(branch(LN_branch(x)));  x = x + .5 drop(FFN(LN_ff(x)));
#   x = LN_final(x)      with branch = rel-pos attention OR cgMLP, FFNs / LN_mac / LN_ff / LN_final SHARED
# i.e. a single-branch Branchformer block without merge_proj.  Each stream is one autograd node built
# on block_forward / block_backward through a view of the layer that names the stream's modules the
# way a Branchformer block does; the shared parameters get the sum of both nodes' gradients.
# ---------------------------------------------------------------------------------------------------
class _StreamView:
    """The modules one stream of a TailoredEncoderLayer uses, under MyBranchformerEncoderLayer names."""
    use_two_branches = False
    merge_method = "identity"
    attn_branch_drop_rate = 0.0
    merge_proj = torch.nn.Identity()

    def __init__(self, layer, tag: str):
        self.size = layer.size
        self.training = layer.training
        self.dropout = layer.dropout
        self.attn = getattr(layer, tag + "_attn")
        self.cgmlp = getattr(layer, tag + "_cgmlp")
        self.norm_mha = getattr(layer, tag + "_norm_mha", None)
        self.norm_mlp = getattr(layer, tag + "_norm_cgmlp", None)
        for n in ("norm_ff_macaron", "feed_forward_macaron", "norm_ff", "feed_forward", "norm_final"):
            setattr(self, n, getattr(layer, n))
        self.rename = {"attn.": tag + "_attn.", "cgmlp.": tag + "_cgmlp.", "norm_mha.": tag + "_norm_mha.",
                       "norm_mlp.": tag + "_norm_cgmlp."}

    def real_name(self, key: str) -> str:
        for k, v in self.rename.items():
            if key.startswith(k):
                return v + key[len(k):]
        return key


class _TailoredStreamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, view, aux, names, x, *params):
        y, sv = block_forward(view, aux, x.contiguous())
        ctx.view, ctx.aux, ctx.sv, ctx.names, ctx.params = view, aux, sv, names, params
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, g = block_backward(ctx.view, ctx.aux, ctx.sv, dy.contiguous())
        ctx.sv = None
        g = {ctx.view.real_name(k): v for k, v in g.items()}
        grads = []
        for n, p in zip(ctx.names, ctx.params):
            gp = g.get(n)
            grads.append(gp.reshape(p.shape) if gp is not None and p.requires_grad else None)
        return (None, None, None, dx) + tuple(grads)


def tailored_layer_forward(layer, a2d, v2d, B: int, T: int, lens_a, lens_v, pos_a, pos_v):
    """Training forward of one TailoredEncoderLayer on (B*T, d) streams.  One stochastic-depth draw
    for the layer (:152-165), then the audio stream's dropout sites, then the video stream's."""
    layer._check_supported()
    stoch = 1.0
    if layer.training and layer.stochastic_depth_rate > 0:
        skip = torch.rand(1).item() < layer.stochastic_depth_rate
        stoch = 1.0 / (1 - layer.stochastic_depth_rate)
        if skip:
            return a2d, v2d
    names, params = _named(layer)
    outs = []
    for tag, x, lens, pos in (("acoustic", a2d, lens_a, pos_a), ("visual", v2d, lens_v, pos_v)):
        view = _StreamView(layer, tag)
        pos2d, pos2d_T = pos if pos is not None else (None, None)
        aux = SimpleNamespace(B=B, T=T, lens=lens, pos2d=pos2d, pos2d_T=pos2d_T, stoch=stoch,
                              drop_attn=False, drop=draw_block_masks(view, B, T, x.device))
        outs.append(_TailoredStreamFn.apply(view, aux, names, x, *params))
    return outs[0], outs[1]


def _av_tap(enc, norm_a, norm_v, a, v, la, lv, B, T, ctc, fusion, inter, idx):
    """One audio-visual InterCTC tap (tailored/encoder.py:270-318, conventional/encoder.py:156-199):
    both streams normalised, fused (the fused stream is what the intermediate CTC loss sees), and with
    self-conditioning the posteriors - of the fused stream, or of each stream - are added back to
    BOTH streams through conditioning_layer.  Returns the (possibly conditioned) streams."""
    d = a.shape[1]
    ta = _LayerNormFn.apply(a, norm_a.weight, norm_a.bias, 1e-12, 1.0) if norm_a is not None else a
    tv = _LayerNormFn.apply(v, norm_v.weight, norm_v.bias, 1e-12, 1.0) if norm_v is not None else v
    fused = fusion_forward(fusion, ta, tv, la, lv, B, T)
    inter.append((idx, fused.view(B, T, d)))
    if enc.interctc_use_conditioning:
        cl = enc.conditioning_layer
        wc, bc = ctc.ctc_lo.weight, ctc.ctc_lo.bias
        if enc.audiovisual_interctc_conditioning:
            a = _CondFn.apply(a, fused, wc, bc, cl.weight, cl.bias)
            v = _CondFn.apply(v, fused, wc, bc, cl.weight, cl.bias)
        else:
            a = _CondFn.apply(a, ta, wc, bc, cl.weight, cl.bias)
            v = _CondFn.apply(v, tv, wc, bc, cl.weight, cl.bias)
    return a, v


def _check_av_taps(enc, ctc, fusion):
    if fusion is None or not hasattr(fusion, "run"):
        raise ValueError("audio-visual InterCTC taps need the B200 `audiovisual_fusion` module")
    if enc.interctc_use_conditioning and (ctc is None or enc.conditioning_layer is None):
        raise ValueError("InterCTC self-conditioning needs the `ctc` module and an assigned "
                         "`conditioning_layer`")


def conventional_encoder_forward(enc, audio_pad, audio_masks, video_pad, video_masks, ctc=None,
                                 fusion=None):
    """Training forward of ConventionalEncoder (conventional/encoder.py:116-215): the two stacks, or
    with InterCTC taps their layer-zipped form."""
    if len(enc.interctc_layer_idx) == 0:
        ya, _, _ = encoder_forward(enc.acoustic_encoder, audio_pad, None, masks=audio_masks)
        yv, _, _ = encoder_forward(enc.visual_encoder, video_pad, None, masks=video_masks)
        return ya, audio_masks, yv, video_masks, None
    _check_av_taps(enc, ctc, fusion)
    ea, ev = enc.acoustic_encoder, enc.visual_encoder
    if ea.embed is not None or ev.embed is not None:
        raise NotImplementedError("the conventional AV encoder expects (x, pos_emb) inputs")
    (audio, a_pos), (video, v_pos) = audio_pad, video_pad
    B, T, d = audio.shape
    dev = audio.device
    a = audio.reshape(B * T, d).contiguous().float()
    v = video.reshape(B * T, d).contiguous().float()
    la = engine.lens_from_mask(audio_masks, B, T, dev)
    lv = engine.lens_from_mask(video_masks, B, T, dev)

    def pos_pair(pos):
        p2 = pos.reshape(-1, d).contiguous().float()
        return p2, ob.transpose_2d(p2, pad=True)

    pa, pv = pos_pair(a_pos), pos_pair(v_pos)
    inter = []
    for i, (layer_a, layer_v) in enumerate(zip(ea.encoders, ev.encoders)):
        a = run_block(layer_a, a, B, T, la, a_pos, shared_pos=pa)
        v = run_block(layer_v, v, B, T, lv, v_pos, shared_pos=pv)
        if (i + 1) in enc.interctc_layer_idx:
            a, v = _av_tap(enc, ea.after_norm if ea.normalize_before else None,
                           ev.after_norm if ev.normalize_before else None, a, v, la, lv, B, T, ctc,
                           fusion, inter, i + 1)
    if ea.normalize_before:
        a = _LayerNormFn.apply(a, ea.after_norm.weight, ea.after_norm.bias, 1e-12, 1.0)
    if ev.normalize_before:
        v = _LayerNormFn.apply(v, ev.after_norm.weight, ev.after_norm.bias, 1e-12, 1.0)
    return (a.view(B, T, d), inter), audio_masks, v.view(B, T, d), video_masks, None


def tailored_encoder_forward(enc, audio_pad, audio_masks, video_pad, video_masks, ctc=None, fusion=None):
    """Training forward of TailoredEncoder (tailored/encoder.py:221-330): modality encoding, the layer
    stack with optional audio-visual InterCTC taps, the shared after_norm on both streams."""
    taps_on = len(enc.interctc_layer_idx) > 0
    if taps_on:
        _check_av_taps(enc, ctc, fusion)
    audio, a_pos = audio_pad if isinstance(audio_pad, tuple) else (audio_pad, None)
    video, v_pos = video_pad if isinstance(video_pad, tuple) else (video_pad, None)
    if audio.shape != video.shape:
        raise NotImplementedError("the B200 tailored encoder expects time-aligned streams of equal shape")
    B, T, d = audio.shape
    M = B * T
    dev = audio.device
    me = enc.modality_encoding.weight
    # modality encoding (:251-263): a broadcast add, left to torch (and its autograd) like the
    # inference path does
    a = (audio.float() + me[0]).reshape(M, d).contiguous()
    v = (video.float() + me[1]).reshape(M, d).contiguous()
    la = engine.lens_from_mask(audio_masks, B, T, dev)
    lv = engine.lens_from_mask(video_masks, B, T, dev)

    def pos_pair(pos):
        if pos is None:
            return None
        p2 = pos.reshape(-1, d).contiguous().float()
        return p2, ob.transpose_2d(p2, pad=True)

    pa, pv = pos_pair(a_pos), pos_pair(v_pos)
    inter = []
    norm = enc.after_norm if enc.normalize_before else None
    for i, layer in enumerate(enc.encoders):
        a, v = tailored_layer_forward(layer, a, v, B, T, la, lv, pa, pv)
        if taps_on and (i + 1) in enc.interctc_layer_idx:
            a, v = _av_tap(enc, norm, norm, a, v, la, lv, B, T, ctc, fusion, inter, i + 1)
    if enc.normalize_before:
        a = _LayerNormFn.apply(a, enc.after_norm.weight, enc.after_norm.bias, 1e-12, 1.0)
        v = _LayerNormFn.apply(v, enc.after_norm.weight, enc.after_norm.bias, 1e-12, 1.0)
    a_out = a.view(B, T, d)
    return ((a_out, inter) if inter else a_out), audio_masks, v.view(B, T, d), video_masks, None


# ---------------------------------------------------------------------------------------------------
# AdaptiveAudioVisualFusion (src/audiovisual_fusion/adaptive_audiovisual_fusion.py:113-215)
#   (w_a, w_v) = learned_ave(audio, video) with one mask per modality   (or fixed weights)
#   out = norm_final(FFN(w_a audio + w_v video)),  FFN = w_2(dropout(act(w_1 .)))   - no residual
# ---------------------------------------------------------------------------------------------------
def fusion_forward_impl(Fm, aux, a: torch.Tensor, v: torch.Tensor):
    B, T = aux.B, aux.T
    d = Fm.input_size
    dev = a.device
    ff = Fm.audiovisual_layer
    learned = Fm.merge_method == "learned_ave" and not aux.drop_acoustic
    if Fm.merge_method == "learned_ave":
        if learned:
            ap, aw = Fm.acoustic_pooling_proj, Fm.acoustic_weight_proj
            vp, vw = Fm.visual_pooling_proj, Fm.visual_weight_proj
            d1, d2 = ops.row_dots(a, ap.weight.reshape(-1), aw.weight.reshape(-1),
                                  v, vp.weight.reshape(-1), vw.weight.reshape(-1))
            scal = torch.cat([ap.bias, vp.bias, aw.bias, vw.bias]).float()
            w_a, w_v = ops.merge_weights_dev(d1, d2, aux.lens_a, scal, d, B, T, lens2=aux.lens_v)
        else:   # acoustic branch dropped for this step (:138-143)
            w_a = torch.zeros((B,), device=dev, dtype=F32)
            w_v = torch.ones((B,), device=dev, dtype=F32)
        Fm.acoustic_weight, Fm.visual_weight = w_a.detach().view(B, 1, 1), w_v.detach().view(B, 1, 1)
    else:
        wa = float(Fm.acoustic_weight)
        w_a = torch.full((B,), wa, device=dev, dtype=F32)
        w_v = torch.full((B,), 1.0 - wa, device=dev, dtype=F32)
    m = ops.scale_add_rows(a, v, w_a, w_v, T)
    act = engine.act_code(ff.activation_type)
    z = ops.gemm_bias_act(m, ff.w_1.weight, ff.w_1.bias)
    h = ob.act_fwd(z, act, mask=aux.mask_h)
    y0 = ops.gemm_bias_act(h, ff.w_2.weight, ff.w_2.bias)
    out = ops.layernorm(y0, Fm.norm_final.weight, Fm.norm_final.bias, eps=1e-12)
    return out, dict(a=a, v=v, w_a=w_a, w_v=w_v, m=m, z=z, y0=y0, learned=learned)


def fusion_backward_impl(Fm, aux, sv: dict, dout: torch.Tensor):
    B, T = aux.B, aux.T
    d = Fm.input_size
    g: Dict[str, torch.Tensor] = {}
    ff = Fm.audiovisual_layer
    act = engine.act_code(ff.activation_type)
    dy0 = _ln_bwd(g, sv["y0"], Fm.norm_final, dout, "norm_final")
    hT = ob.act_fwd_t(sv["z"], act, mask=aux.mask_h)
    dh = _lin_bwd(g, None, ff.w_2.weight, dy0, "audiovisual_layer.w_2.weight", "audiovisual_layer.w_2.bias",
                  xT=hT)
    dz, dzT = ob.act_bwd_t(sv["z"], _mul(dh, aux.mask_h), act)
    dm = _lin_bwd(g, sv["m"], ff.w_1.weight, dz, "audiovisual_layer.w_1.weight",
                  "audiovisual_layer.w_1.bias", dyT=dzT)
    if sv["learned"]:
        ap, aw = Fm.acoustic_pooling_proj, Fm.acoustic_weight_proj
        vp, vw = Fm.visual_pooling_proj, Fm.visual_weight_proj
        scal = torch.cat([ap.bias, aw.bias, vp.bias, vw.bias]).float()
        da, dv, gr = ob.merge_learned_ave_bwd(sv["a"], sv["v"], dm, aux.lens_a, ap.weight.reshape(-1),
                                              aw.weight.reshape(-1), vp.weight.reshape(-1),
                                              vw.weight.reshape(-1), scal, B, T, lens2=aux.lens_v)
        for k, name in enumerate(("acoustic_pooling_proj", "acoustic_weight_proj", "visual_pooling_proj",
                                  "visual_weight_proj")):
            _acc(g, name + ".weight", gr[k * d:(k + 1) * d].reshape(1, d))
            _acc(g, name + ".bias", gr[4 * d + k:4 * d + k + 1])
    else:
        zero = torch.zeros((B,), device=dout.device, dtype=F32)
        da = ops.scale_add_rows(dm, dm, sv["w_a"], zero, T)
        dv = ops.scale_add_rows(dm, dm, sv["w_v"], zero, T)
    return da, dv, g


class _FusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Fm, aux, names, a, v, *params):
        out, sv = fusion_forward_impl(Fm, aux, a.contiguous(), v.contiguous())
        ctx.Fm, ctx.aux, ctx.sv, ctx.names, ctx.params = Fm, aux, sv, names, params
        return out

    @staticmethod
    def backward(ctx, dout):
        da, dv, g = fusion_backward_impl(ctx.Fm, ctx.aux, ctx.sv, dout.contiguous())
        ctx.sv = None
        grads = []
        for n, p in zip(ctx.names, ctx.params):
            gp = g.get(n)
            grads.append(gp.reshape(p.shape) if gp is not None and p.requires_grad else None)
        return (None, None, None, da, dv) + tuple(grads)


def fusion_forward(Fm, a2d: torch.Tensor, v2d: torch.Tensor, lens_a: torch.Tensor, lens_v: torch.Tensor,
                   B: int, T: int) -> torch.Tensor:
    """Training forward of AdaptiveAudioVisualFusion on (B*T, d) streams as one autograd node.
    Randomness in the reference's order: the acoustic-branch drop draws torch.rand(1) on the host
    (:138-143), then the FFN's hidden dropout."""
    if Fm.input_size != 256 or Fm._output_size != 256:
        raise NotImplementedError("the B200 fusion is built for size=256")
    drop_acoustic = False
    if Fm.merge_method == "learned_ave" and Fm.training and Fm.acoustic_branch_drop_rate > 0:
        drop_acoustic = torch.rand(1).item() < Fm.acoustic_branch_drop_rate
    mask_h = None
    if Fm.training:
        ff = Fm.audiovisual_layer
        m = draw_mask((B, T, ff.w_1.out_features), float(ff.dropout_rate), a2d.device)
        mask_h = None if m is None else m.view(B * T, -1)
    aux = SimpleNamespace(B=B, T=T, lens_a=lens_a, lens_v=lens_v, drop_acoustic=drop_acoustic,
                          mask_h=mask_h)
    names, params = _named(Fm)
    return _FusionFn.apply(Fm, aux, names, a2d, v2d, *params)


def make_aux(layer, B: int, T: int, lens: torch.Tensor, pos_emb: Optional[torch.Tensor]):
    """Per-call constants of a block plus the training-only random decisions, drawn from the host
    RNG exactly where the reference draws them (encoder_layer.py:176-189, 233-239).  Returns None
    when stochastic depth skips the layer."""
    stoch = 1.0
    if layer.training and layer.stochastic_depth_rate > 0:
        skip = torch.rand(1).item() < layer.stochastic_depth_rate
        stoch = 1.0 / (1 - layer.stochastic_depth_rate)
        if skip:
            return None
    drop_attn = False
    if (layer.use_two_branches and layer.merge_method == "learned_ave" and layer.training
            and layer.attn_branch_drop_rate > 0):
        drop_attn = torch.rand(1).item() < layer.attn_branch_drop_rate
    pos2d = pos2d_T = None
    if pos_emb is not None:
        pos2d = pos_emb.reshape(-1, pos_emb.shape[-1]).contiguous().float()
        pos2d_T = ob.transpose_2d(pos2d, pad=True)
    return SimpleNamespace(B=B, T=T, lens=lens, pos2d=pos2d, pos2d_T=pos2d_T, stoch=stoch,
                           drop_attn=drop_attn, drop=draw_block_masks(layer, B, T, lens.device))


def draw_block_masks(layer, B: int, T: int, device) -> Dict[str, object]:
    """The dropout masks of one block, drawn in the reference's call order (module docstring).
    Empty in eval mode or when every rate is 0."""
    dm: Dict[str, object] = {}
    if not layer.training:
        return dm
    d = layer.size
    p_out = float(layer.dropout.p)

    def site(key, shape, p):
        m = draw_mask(shape, p, device)        # drawn with the reference's (B, T, C) shape,
        if m is not None:                      # used on the (B*T, C) activations
            dm[key] = m.view(B * T, shape[-1])

    def ffn(tag, ff):
        site(tag + "_h", (B, T, ff.w_1.out_features), float(ff.dropout_rate))
        site(tag + "_o", (B, T, d), p_out)

    ffn("ffm", layer.feed_forward_macaron)
    if layer.attn is not None:
        p_att = float(layer.attn.dropout_rate)
        m = draw_mask((B, layer.attn.h, T, T), p_att, device)
        if m is not None:
            dm["att"] = (_keep_bytes(m, T), 1.0 / (1.0 - p_att))
        site("x1", (B, T, d), p_out)
    if layer.cgmlp is not None:
        site("csgu", (B, T, layer.cgmlp.channel_proj2.in_features), float(layer.cgmlp.csgu.dropout_rate))
        site("x2", (B, T, d), p_out)
    if not isinstance(layer.merge_proj, torch.nn.Identity):
        site("merge", (B, T, d), p_out)
    ffn("ff", layer.feed_forward)
    return dm


def run_block(layer, x2d: torch.Tensor, B: int, T: int, lens: torch.Tensor,
              pos_emb: Optional[torch.Tensor], shared_pos=None) -> torch.Tensor:
    """Training forward of one block on (B*T, d) activations, as an autograd node."""
    layer._check_supported()
    aux = make_aux(layer, B, T, lens, pos_emb if shared_pos is None else None)
    if aux is None:
        return x2d                                     # stochastic depth: the layer is skipped
    if shared_pos is not None:
        aux.pos2d, aux.pos2d_T = shared_pos
    names, params = _named(layer)
    return _BlockFn.apply(layer, aux, names, x2d, *params)


def encoder_forward(enc, xs_pad, ilens, max_layer=None, masks=None, ctc=None):
    """Training forward of MyBranchformerEncoder (encoder.py:324-412: the layer loop, the max_layer
    early exit, InterCTC taps with optional self-conditioning).  Returns (out (B,T,d) or (out,
    [(layer, tap (B,T,d)), ...]), olens, None) with an autograd graph behind every returned tensor.
    `masks` (B,1,T): given by the AV wrappers instead of `ilens`."""
    taps_on = len(enc.interctc_layer_idx) > 0
    if taps_on and enc.interctc_use_conditioning and (ctc is None or enc.conditioning_layer is None):
        raise ValueError("InterCTC self-conditioning needs the `ctc` module and an assigned "
                         "`conditioning_layer` (espnet_model.py:106-112)")
    x_in = xs_pad[0] if isinstance(xs_pad, tuple) else xs_pad
    dev = x_in.device
    d = enc._output_size
    Tin = x_in.size(1)
    if masks is None:
        masks = (torch.arange(Tin, device=dev)[None, :] < ilens.to(dev)[:, None]).unsqueeze(1)
    if enc.embed is None:
        xs, pos_emb = xs_pad
        B, T, _ = xs.shape
        x = xs.reshape(B * T, d).contiguous().float()
    elif isinstance(enc.embed, torch.nn.Sequential):          # input_layer == "linear"
        # Sequential(Linear, LayerNorm, Dropout(dropout_rate), RelPositionalEncoding): dropout, then
        # x * sqrt(d), then the positional-encoding module's dropout on x and on pos_emb
        # (encoder.py:125-128; espnet embedding.py RelPositionalEncoding.forward)
        B, T, Fd = x_in.shape
        lin, ln, pe = enc.embed[0], enc.embed[1], enc.embed[3]
        e0 = _LinearFn.apply(x_in.reshape(B * T, Fd).contiguous().float(), lin.weight, lin.bias)
        x = _LayerNormFn.apply(e0, ln.weight, ln.bias, ln.eps, math.sqrt(d))
        pos_emb = pe.pos_emb(T, dev)
        if enc.training:
            for shape, p in (((B, T, d), float(enc.embed[2].p)), ((B, T, d), float(pe.dropout_rate))):
                m = draw_mask(shape, p, dev)
                if m is not None:
                    x = _MaskFn.apply(x, m.view(B * T, d))
            m = draw_mask(tuple(pos_emb.shape), float(pe.dropout_rate), dev)
            if m is not None:
                pos_emb = _mul(pos_emb.reshape(-1, d).contiguous().float(), m.view(-1, d)).view(pos_emb.shape)
    else:                                                      # input_layer == "conv2d"
        B, Tin, _ = x_in.shape
        T = ((Tin - 1) // 2 - 1) // 2
        pe = enc.embed.out[1]
        x = _ScaleFn.apply(conv2d_front_forward(enc.embed, x_in), math.sqrt(d))
        pos_emb = pe.pos_emb(T, dev)
        x, pos_emb = posenc_dropouts(pe, x, pos_emb, B, T, enc.training)
        masks = masks[:, :, :-2:2][:, :, :-2:2]
    lens = masks.reshape(B, -1).sum(dim=1).to(torch.int32)
    pos2d = pos_emb.reshape(-1, d).contiguous().float()
    shared = (pos2d, ob.transpose_2d(pos2d, pad=True))
    n = len(enc.encoders)
    last = n - 1 if (taps_on or max_layer is None or not 0 <= max_layer < n) else max_layer
    inter = []
    for i, layer in enumerate(enc.encoders):
        if i > last:
            break
        x = run_block(layer, x, B, T, lens, pos_emb, shared_pos=shared)
        if taps_on and (i + 1) in enc.interctc_layer_idx:
            # intermediate outputs are normalised too (:387-389), then fed back as posteriors (:393-401)
            tap = (_LayerNormFn.apply(x, enc.after_norm.weight, enc.after_norm.bias, 1e-12, 1.0)
                   if enc.normalize_before else x)
            inter.append((i + 1, tap.view(B, T, d)))
            if enc.interctc_use_conditioning:
                cl = enc.conditioning_layer
                x = _CondFn.apply(x, tap, ctc.ctc_lo.weight, ctc.ctc_lo.bias, cl.weight, cl.bias)
    if enc.normalize_before:
        x = _LayerNormFn.apply(x, enc.after_norm.weight, enc.after_norm.bias, 1e-12, 1.0)
    olens = masks.squeeze(1).sum(1)
    if inter:
        return (x.view(B, T, d), inter), olens, None
    return x.view(B, T, d), olens, None
