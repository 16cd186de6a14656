"""TEST INFRASTRUCTURE: the backward of one two-branch learned_ave Branchformer block composed BY
HAND from the leaf formulas of oracle/bwd_formulas.py, in the order and with the saved tensors the
CUDA training path uses (tailored_avsr_b200/training.py): which activations are kept from the forward,
where the residual gradients join, how the merge gradient splits into the two branches.  Verified
against torch.autograd on ref_path.branchformer_layer (tests/test_bwd_formulas_cpu.py).

Forward being differentiated (src/encoder/branchformer/encoder_layer.py:191-316, eval arithmetic):
    x_a = x + 0.5 FFN_mac(LN_ffmac(x))
    x1  = Wo . attn(LN_mha(x_a)) + bo          x2 = W2 . csgu(gelu(Wc1 LN_mlp(x_a) + bc1)) + b2
    m   = w1 x1 + w2 x2 (learned_ave)          x_b = x_a + Wm m + bm
    y0  = x_b + 0.5 FFN(LN_ff(x_b))            y = LN_final(y0)
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import bwd_formulas as bw
from . import ref_path

SD = Dict[str, torch.Tensor]


def _ln(x, sd, name, eps=1e-12):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _acc(grads, key, val):
    grads[key] = grads[key] + val if key in grads else val


def _ffn_fwd(xn, sd, p):
    z = F.linear(xn, sd[p + ".w_1.weight"], sd[p + ".w_1.bias"])
    hdn = z * torch.sigmoid(z)
    return z, hdn, F.linear(hdn, sd[p + ".w_2.weight"], sd[p + ".w_2.bias"])


def _ffn_bwd(xn, z, hdn, dout, sd, p, grads):
    """dout = gradient w.r.t. the FFN output (already scaled by 0.5).  Returns d xn."""
    dh, dw2, db2 = bw.linear_bwd(hdn, sd[p + ".w_2.weight"], dout)
    dz = bw.swish_bwd(z, dh)
    dxn, dw1, db1 = bw.linear_bwd(xn, sd[p + ".w_1.weight"], dz)
    for k, v in ((".w_2.weight", dw2), (".w_2.bias", db2), (".w_1.weight", dw1), (".w_1.bias", db1)):
        _acc(grads, p + k, v)
    return dxn


def _ln_bwd(x, dy, sd, name, grads, eps=1e-12):
    dx, dg, db = bw.layernorm_bwd(x, sd[name + ".weight"], dy, eps)
    _acc(grads, name + ".weight", dg)
    _acc(grads, name + ".bias", db)
    return dx


def layer_forward_backward(x, pos_emb, mask, sd: SD, prefix: str, dy, heads: int = 4):
    """Returns (y, dx, grads) for a two-branch learned_ave block; grads maps parameter names to
    gradients.  Everything that the backward reads is computed in the forward part below and named
    like the tensors the CUDA training forward will save."""
    B, T, D = x.shape
    dk = D // heads
    lens = mask.squeeze(1).sum(1)
    P = prefix
    # ------------------------------- forward (saving) ----------------------------------------
    xn0 = _ln(x, sd, P + ".norm_ff_macaron")
    z1, h1, f1 = _ffn_fwd(xn0, sd, P + ".feed_forward_macaron")
    x_a = x + 0.5 * f1
    xa = _ln(x_a, sd, P + ".norm_mha")
    xm = _ln(x_a, sd, P + ".norm_mlp")
    A = P + ".attn"
    proj = lambda t, n: F.linear(t, sd[A + f".linear_{n}.weight"], sd[A + f".linear_{n}.bias"])  # noqa: E731
    q = proj(xa, "q").view(B, T, heads, dk).transpose(1, 2)
    k = proj(xa, "k").view(B, T, heads, dk).transpose(1, 2)
    v = proj(xa, "v").view(B, T, heads, dk).transpose(1, 2)
    pp = F.linear(pos_emb[0], sd[A + ".linear_pos.weight"]).view(2 * T - 1, heads, dk).transpose(0, 1)
    u, vb = sd[A + ".pos_bias_u"], sd[A + ".pos_bias_v"]
    # attention core forward (same arithmetic as ref_path.rel_pos_mha)
    qu, qv = q + u[None, :, None, :], q + vb[None, :, None, :]
    idx = (T - 1 - torch.arange(T).unsqueeze(1) + torch.arange(T).unsqueeze(0))
    s = (qu @ k.transpose(-2, -1) + (qv @ pp.transpose(-2, -1)[None]).gather(-1, idx.expand(B, heads, T, T)))
    s = s / math.sqrt(dk)
    inv = mask.unsqueeze(1).eq(0)
    s = s.masked_fill(inv, torch.finfo(s.dtype).min)
    attn = torch.softmax(s, dim=-1).masked_fill(inv, 0.0)
    ctx = (attn @ v).transpose(1, 2).reshape(B, T, D)
    x1 = F.linear(ctx, sd[A + ".linear_out.weight"], sd[A + ".linear_out.bias"])
    C = P + ".cgmlp"
    zc = F.linear(xm, sd[C + ".channel_proj1.0.weight"], sd[C + ".channel_proj1.0.bias"])
    hc = F.gelu(zc)
    Ch = hc.shape[-1] // 2
    r_, g_ = hc[..., :Ch], hc[..., Ch:]
    gn = _ln(g_, sd, C + ".csgu.norm")
    cw, cbias = sd[C + ".csgu.conv.weight"], sd[C + ".csgu.conv.bias"]
    kk = cw.shape[-1]
    conv = F.conv1d(gn.transpose(1, 2), cw, cbias, padding=(kk - 1) // 2, groups=Ch).transpose(1, 2)
    uu = r_ * conv
    x2 = F.linear(uu, sd[C + ".channel_proj2.weight"], sd[C + ".channel_proj2.bias"])
    om1 = ref_path._pool_weight(x1, mask, sd, P + ".pooling_proj1", P + ".weight_proj1")
    om2 = ref_path._pool_weight(x2, mask, sd, P + ".pooling_proj2", P + ".weight_proj2")
    w = torch.softmax(torch.cat([om1, om2], dim=-1), dim=-1)
    m = w[:, 0, None, None] * x1 + w[:, 1, None, None] * x2
    x_b = x_a + F.linear(m, sd[P + ".merge_proj.weight"], sd[P + ".merge_proj.bias"])
    xf = _ln(x_b, sd, P + ".norm_ff")
    z2, h2, f2 = _ffn_fwd(xf, sd, P + ".feed_forward")
    y0 = x_b + 0.5 * f2
    y = _ln(y0, sd, P + ".norm_final")
    # ------------------------------- backward -------------------------------------------------
    g: Dict[str, torch.Tensor] = {}
    dy0 = _ln_bwd(y0, dy, sd, P + ".norm_final", g)
    dxf = _ffn_bwd(xf, z2, h2, 0.5 * dy0, sd, P + ".feed_forward", g)
    dx_b = dy0 + _ln_bwd(x_b, dxf, sd, P + ".norm_ff", g)                 # residual joins here
    dm, dwm, dbm = bw.linear_bwd(m, sd[P + ".merge_proj.weight"], dx_b)
    _acc(g, P + ".merge_proj.weight", dwm)
    _acc(g, P + ".merge_proj.bias", dbm)
    flat = lambda t: t.reshape(-1) if t.numel() > 1 else t.reshape(())     # noqa: E731
    outs = bw.learned_ave_merge_bwd(
        x1, x2, lens,
        flat(sd[P + ".pooling_proj1.weight"]), flat(sd[P + ".pooling_proj1.bias"]),
        flat(sd[P + ".weight_proj1.weight"]), flat(sd[P + ".weight_proj1.bias"]),
        flat(sd[P + ".pooling_proj2.weight"]), flat(sd[P + ".pooling_proj2.bias"]),
        flat(sd[P + ".weight_proj2.weight"]), flat(sd[P + ".weight_proj2.bias"]), (dm,))
    (dx1, da1, dc1, db1_, de1), (dx2, da2, dc2, db2_, de2) = outs
    for tag, da, dc, db_, de in (("1", da1, dc1, db1_, de1), ("2", da2, dc2, db2_, de2)):
        _acc(g, P + f".pooling_proj{tag}.weight", da.reshape(1, -1))
        _acc(g, P + f".pooling_proj{tag}.bias", dc.reshape(1))
        _acc(g, P + f".weight_proj{tag}.weight", db_.reshape(1, -1))
        _acc(g, P + f".weight_proj{tag}.bias", de.reshape(1))
    # ---- cgMLP branch ----
    du_, dw2c, db2c = bw.linear_bwd(uu, sd[C + ".channel_proj2.weight"], dx2)
    _acc(g, C + ".channel_proj2.weight", dw2c)
    _acc(g, C + ".channel_proj2.bias", db2c)
    dhc, dng, dnb, dcw, dcb = bw.csgu_bwd(hc, sd[C + ".csgu.norm.weight"], sd[C + ".csgu.norm.bias"],
                                          cw, cbias, du_)
    for kname, val in ((".csgu.norm.weight", dng), (".csgu.norm.bias", dnb), (".csgu.conv.weight", dcw),
                       (".csgu.conv.bias", dcb)):
        _acc(g, C + kname, val)
    dzc = bw.gelu_bwd(zc, dhc)
    dxm, dwc1, dbc1 = bw.linear_bwd(xm, sd[C + ".channel_proj1.0.weight"], dzc)
    _acc(g, C + ".channel_proj1.0.weight", dwc1)
    _acc(g, C + ".channel_proj1.0.bias", dbc1)
    # ---- attention branch ----
    dctx, dwo, dbo = bw.linear_bwd(ctx, sd[A + ".linear_out.weight"], dx1)
    _acc(g, A + ".linear_out.weight", dwo)
    _acc(g, A + ".linear_out.bias", dbo)
    do = dctx.view(B, T, heads, dk).transpose(1, 2)
    dq, dkk, dv, dp, du_b, dv_b = bw.relpos_attn_core_bwd(q, k, v, pp, u, vb, lens, do)
    _acc(g, A + ".pos_bias_u", du_b)
    _acc(g, A + ".pos_bias_v", dv_b)
    _acc(g, A + ".linear_pos.weight", dp.transpose(0, 1).reshape(2 * T - 1, D).t() @ pos_emb[0])
    dxa = torch.zeros_like(xa)
    for name, dt in (("q", dq), ("k", dkk), ("v", dv)):
        d2 = dt.transpose(1, 2).reshape(B, T, D)
        dxi, dwi, dbi = bw.linear_bwd(xa, sd[A + f".linear_{name}.weight"], d2)
        dxa = dxa + dxi
        _acc(g, A + f".linear_{name}.weight", dwi)
        _acc(g, A + f".linear_{name}.bias", dbi)
    # ---- the two branch LayerNorms and the merge residual meet at x_a ----
    dx_a = dx_b + _ln_bwd(x_a, dxa, sd, P + ".norm_mha", g) + _ln_bwd(x_a, dxm, sd, P + ".norm_mlp", g)
    dxn0 = _ffn_bwd(xn0, z1, h1, 0.5 * dx_a, sd, P + ".feed_forward_macaron", g)
    dx = dx_a + _ln_bwd(x, dxn0, sd, P + ".norm_ff_macaron", g)
    return y, dx, g


def encoder_forward_backward(xs, ilens, sd: SD, cfg: dict, dout_fn):
    """Whole single-stream encoder (linear front end, two-branch learned_ave blocks, swish FFN):
    forward, then the backward chained by hand through after_norm, the blocks in reverse (each
    block's backward is layer_forward_backward on its saved input) and the embed (Linear +
    LayerNorm(1e-5) + x sqrt(d)).  `dout_fn(out, olens)` returns d loss / d out (the CTC backward,
    which already exists as a CUDA kernel).  Returns (out, grads incl. "input")."""
    assert cfg.get("input_layer") == "linear" and cfg.get("merge_method", "learned_ave") == "learned_ave"
    d = cfg.get("output_size", 256)
    n = cfg.get("num_blocks", 12)
    heads = cfg.get("attention_heads", 4)
    T = xs.shape[1]
    masks = ref_path.make_valid_mask(ilens, T)
    e0 = F.linear(xs, sd["embed.0.weight"], sd["embed.0.bias"])
    e1 = F.layer_norm(e0, (d,), sd["embed.1.weight"], sd["embed.1.bias"], 1e-5)
    x = e1 * math.sqrt(d)
    pos = ref_path.rel_pos_emb(T, d).to(xs.dtype)
    block_in = []
    for l in range(n):
        block_in.append(x)
        x, _ = ref_path.branchformer_layer(x, pos, masks, sd, f"encoders.{l}", heads=heads,
                                           kernel=cfg.get("cgmlp_conv_kernel", 31), act="swish")
    out = _ln(x, sd, "after_norm")
    g: Dict[str, torch.Tensor] = {}
    dx = _ln_bwd(x, dout_fn(out, masks.squeeze(1).sum(1)), sd, "after_norm", g)
    for l in reversed(range(n)):
        _, dx, gl = layer_forward_backward(block_in[l], pos, masks, sd, f"encoders.{l}", dx, heads=heads)
        for k_, v_ in gl.items():
            _acc(g, k_, v_)
    de1 = dx * math.sqrt(d)
    de0 = _ln_bwd(e0, de1, sd, "embed.1", g, eps=1e-5)
    dxs, dw, db = bw.linear_bwd(xs, sd["embed.0.weight"], de0)
    _acc(g, "embed.0.weight", dw)
    _acc(g, "embed.0.bias", db)
    g["input"] = dxs
    return out, g
