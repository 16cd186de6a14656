"""TEST INFRASTRUCTURE: explicit (autograd-free) backward formulas of the hot-path leaves, written
the way the CUDA backward kernels compute them (csrc/backward.cu, csrc/attention_bwd.cu) and
verified against torch.autograd on the oracle port (tests/test_bwd_formulas_cpu.py).  This file
fixes the kernels' arithmetic - in particular the rel-shift scatter of the attention backward and
the two-level softmax of the learned_ave merge - and is what their GPU tests compare against.
Plain torch in whatever dtype the caller passes (tests use float64).

Reference forward definitions: espnet leaves as restated in oracle/ref_path.py (SURVEY.md
Appendix A), called from src/encoder/branchformer/encoder_layer.py:193-316."""
from __future__ import annotations

import math
from typing import Tuple

import torch


def layernorm_bwd(x, gamma, dy, eps: float = 1e-12):
    """y = (x - mu) * rstd * gamma + beta over the last axis.  Returns (dx, dgamma, dbeta)."""
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(((x - mu) ** 2).mean(-1, keepdim=True) + eps)
    xh = (x - mu) * rstd
    g = dy * gamma
    dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    red = tuple(range(x.dim() - 1))
    return dx, (dy * xh).sum(red), dy.sum(red)


def swish_bwd(z, dh):
    s = torch.sigmoid(z)
    return dh * s * (1.0 + z * (1.0 - s))


def gelu_bwd(z, dh):
    """exact-erf GELU: d/dz [z Phi(z)] = Phi(z) + z phi(z)."""
    phi = torch.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
    Phi = 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0)))
    return dh * (Phi + z * phi)


def linear_bwd(x, w, dy):
    """y = x w^T + b.  Returns (dx, dw, db); the two GEMMs of the backward (dgrad, wgrad)."""
    x2, dy2 = x.reshape(-1, x.shape[-1]), dy.reshape(-1, dy.shape[-1])
    return dy @ w, dy2.t() @ x2, dy2.sum(0)


def csgu_bwd(h, norm_g, norm_b, conv_w, conv_b, du, eps: float = 1e-12):
    """u = r * (dwconv_k(LN(g)) + b) with h = [r | g] (B,T,2C), conv_w (C,1,k), zero padding (k-1)/2
    in time on the NORMALISED sequence.  Returns (dh, dnorm_g, dnorm_b, dconv_w, dconv_b)."""
    B, T, C2 = h.shape
    C = C2 // 2
    k = conv_w.shape[-1]
    half = (k - 1) // 2
    r, g = h[..., :C], h[..., C:]
    mu = g.mean(-1, keepdim=True)
    rstd = torch.rsqrt(((g - mu) ** 2).mean(-1, keepdim=True) + eps)
    gh = (g - mu) * rstd
    n = gh * norm_g + norm_b                                   # (B,T,C)
    w = conv_w.reshape(C, k)
    npad = torch.nn.functional.pad(n, (0, 0, half, half))      # zero padding in time
    c = conv_b + sum(w[:, j] * npad[:, j:j + T, :] for j in range(k))
    dr = du * c
    dc = du * r                                                # (B,T,C)
    dconv_b = dc.sum((0, 1))
    # dW[ch, j] = sum_{b,t} dc[b,t,ch] * n[b, t + j - half, ch]
    dconv_w = torch.stack([(dc * npad[:, j:j + T, :]).sum((0, 1)) for j in range(k)], dim=1)
    # dn[b,t,ch] = sum_j W[ch,j] * dc[b, t - j + half, ch]   (correlation with the flipped taps)
    dcpad = torch.nn.functional.pad(dc, (0, 0, half, half))
    dn = sum(w[:, j] * dcpad[:, (k - 1 - j):(k - 1 - j) + T, :] for j in range(k))
    # LayerNorm backward on the gate half
    gg = dn * norm_g
    dg = rstd * (gg - gg.mean(-1, keepdim=True) - gh * (gg * gh).mean(-1, keepdim=True))
    dnorm_g = (dn * gh).sum((0, 1))
    dnorm_b = dn.sum((0, 1))
    dh = torch.cat([dr, dg], dim=-1)
    return dh, dnorm_g, dnorm_b, dconv_w.reshape(conv_w.shape), dconv_b


def relpos_attn_core_bwd(q, k, v, p, u, vb, lens, do):
    """Core of espnet RelPositionMultiHeadedAttention between the projections:
        ac = (q+u) k^T, raw = (q+vb) p^T, bd[i,j] = raw[i, T-1-i+j], s = (ac+bd)/sqrt(d),
        P = softmax_j(mask(s)) (masked keys exactly 0), o = P v.
    q,k,v,do: (B,h,T,d); p: (h,2T-1,d); u,vb: (h,d); lens: (B,) valid keys.
    Returns (dq, dk, dv, dp, du, dvb).  The forward probabilities are recomputed (flash style)."""
    B, H, T, d = q.shape
    scale = 1.0 / math.sqrt(d)
    qu = q + u[None, :, None, :]
    qv = q + vb[None, :, None, :]
    ac = qu @ k.transpose(-2, -1)
    raw = qv @ p.transpose(-2, -1)[None]                               # (B,h,T,2T-1)
    idx = (T - 1 - torch.arange(T).unsqueeze(1) + torch.arange(T).unsqueeze(0))
    bd = raw.gather(-1, idx.expand(B, H, T, T))
    s = (ac + bd) * scale
    inv = (torch.arange(T)[None, :] >= lens[:, None])[:, None, None, :]  # True = masked key
    s = s.masked_fill(inv, float("-inf"))
    P = torch.softmax(s, dim=-1)
    P = torch.where(inv, torch.zeros_like(P), P)
    P = torch.nan_to_num(P)                                             # lens == 0: all-masked rows
    dv = P.transpose(-2, -1) @ do
    dP = do @ v.transpose(-2, -1)
    ds = P * (dP - (dP * P).sum(-1, keepdim=True)) * scale              # masked entries: P = 0
    dqu = ds @ k
    dk = ds.transpose(-2, -1) @ qu
    # rel-shift backward: scatter-add ds[i,j] into column T-1-i+j of the (T, 2T-1) band
    draw = torch.zeros_like(raw)
    draw.scatter_add_(-1, idx.expand(B, H, T, T), ds)
    dqv = draw @ p[None]
    dp = (draw.transpose(-2, -1) @ qv).sum(0)                           # (h,2T-1,d)
    return dqu + dqv, dk, dv, dp, dqu.sum((0, 2)), dqv.sum((0, 2))


def learned_ave_merge_bwd(x1, x2, lens, a1, c1, b1, e1, a2, c2, b2, e2, dm_or_w: Tuple):
    """learned_ave merge (encoder_layer.py:241-291) up to m = w1 x1 + w2 x2:
        score_i = (x_i . a_i + c_i)/sqrt(D) masked; s_i = softmax_t; pooled_i = sum_t s_it x_it;
        omega_i = pooled_i . b_i + e_i; (w1, w2) = softmax(omega); m = w1 x1 + w2 x2.
    dm_or_w = (dm,): gradient w.r.t. m (B,T,D).  Returns (dx1, dx2, grads of a_i c_i b_i e_i)."""
    (dm,) = dm_or_w
    B, T, D = x1.shape
    valid = (torch.arange(T)[None, :] < lens[:, None])
    rs = 1.0 / math.sqrt(D)

    def fwd(x, a, c):
        sc = (x @ a + c) * rs
        sc = sc.masked_fill(~valid, float("-inf"))
        s = torch.nan_to_num(torch.softmax(sc, dim=-1))
        s = torch.where(valid, s, torch.zeros_like(s))
        return s, (s.unsqueeze(-1) * x).sum(1)

    s1, pool1 = fwd(x1, a1, c1)
    s2, pool2 = fwd(x2, a2, c2)
    om = torch.stack([pool1 @ b1 + e1, pool2 @ b2 + e2], dim=-1)
    w = torch.softmax(om, dim=-1)                                       # (B,2)
    dw = torch.stack([(dm * x1).sum((1, 2)), (dm * x2).sum((1, 2))], dim=-1)
    dom = w * (dw - (w * dw).sum(-1, keepdim=True))
    outs = []
    for i, (x, s, pool, a, b) in enumerate(((x1, s1, pool1, a1, b1), (x2, s2, pool2, a2, b2))):
        dx = w[:, i, None, None] * dm
        dpool = dom[:, i, None] * b[None, :]                            # (B,D)
        db = (dom[:, i, None] * pool).sum(0)
        de = dom[:, i].sum()
        ds = (x * dpool[:, None, :]).sum(-1)                            # (B,T)
        dx = dx + s.unsqueeze(-1) * dpool[:, None, :]
        dsc = s * (ds - (s * ds).sum(-1, keepdim=True)) * rs           # masked frames: s = 0
        dx = dx + dsc.unsqueeze(-1) * a[None, None, :]
        da = (dsc.unsqueeze(-1) * x).sum((0, 1))
        dc = dsc.sum()
        outs.append((dx, da, dc, db, de))
    return outs


def relpos_attn_core_fwd_stats(q, k, v, p, u, vb, lens):
    """Forward of the attention core returning what a flash-style backward keeps: the output o and
    the per-row log-sum-exp L (natural log, of the scaled masked scores)."""
    B, H, T, d = q.shape
    scale = 1.0 / math.sqrt(d)
    qu = q + u[None, :, None, :]
    qv = q + vb[None, :, None, :]
    idx = (T - 1 - torch.arange(T).unsqueeze(1) + torch.arange(T).unsqueeze(0))
    s = (qu @ k.transpose(-2, -1) + (qv @ p.transpose(-2, -1)[None]).gather(-1, idx.expand(B, H, T, T))) * scale
    inv = (torch.arange(T)[None, :] >= lens[:, None])[:, None, None, :]
    s = s.masked_fill(inv, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)                                     # -inf for all-masked rows
    P = torch.nan_to_num(torch.exp(s - lse.unsqueeze(-1)))
    return P @ v, lse


def relpos_attn_core_bwd_tiled(q, k, v, p, u, vb, lens, do, o, lse, tq: int = 4, tk: int = 4):
    """The same gradients as relpos_attn_core_bwd computed the way a tiled (flash-style, tensor-core)
    kernel will: loop over (query tile, key tile) pairs, recompute the scores of the pair from Q, K
    and the BAND of relative positions the pair touches - band column c of the pair (i0, j0) is
    relative-position row rbase + c with rbase = T - tq - i0 + j0, exactly the forward kernel's
    indexing (attention_sm100.cu) - take P = exp(s - L_i) from the saved row log-sum-exp, use
    D_i = sum_d do_i o_i instead of a row reduction over P dP, and accumulate
    dq (per query tile), dk / dv (per key tile) and dp (per band row)."""
    B, H, T, d = q.shape
    scale = 1.0 / math.sqrt(d)
    qu = q + u[None, :, None, :]
    qv = q + vb[None, :, None, :]
    Drow = (do * o).sum(-1)                                              # (B,H,T)
    dqu = torch.zeros_like(q)
    dqv = torch.zeros_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    dp = torch.zeros_like(p)
    nband = tq + tk - 1
    for i0 in range(0, T, tq):
        i1 = min(i0 + tq, T)
        for j0 in range(0, T, tk):
            j1 = min(j0 + tk, T)
            rbase = T - tq - i0 + j0                                     # band column 0 <-> row rbase
            rows = torch.arange(rbase, rbase + nband).clamp(0, 2 * T - 2)
            band = p[:, rows, :]                                         # (H, nband, d)
            S = qu[:, :, i0:i1] @ k[:, :, j0:j1].transpose(-2, -1)       # (B,H,ti,tj)
            R = qv[:, :, i0:i1] @ band.transpose(-2, -1)[None]           # (B,H,ti,nband)
            # rel-shift inside the pair: key jj of query ii sits at band column (tq - 1 - ii) + jj
            ii = torch.arange(i1 - i0).unsqueeze(1)
            jj = torch.arange(j1 - j0).unsqueeze(0)
            col = (tq - 1 - ii + jj)                                     # (ti,tj)
            s = (S + R.gather(-1, col.expand(B, H, i1 - i0, j1 - j0))) * scale
            keys = torch.arange(j0, j1)
            masked = (keys[None, :] >= lens[:, None])[:, None, None, :]
            L = lse[:, :, i0:i1].unsqueeze(-1)
            Pt = torch.where(masked | torch.isinf(L), torch.zeros_like(s), torch.exp(s - L))
            dv[:, :, j0:j1] += Pt.transpose(-2, -1) @ do[:, :, i0:i1]
            dP = do[:, :, i0:i1] @ v[:, :, j0:j1].transpose(-2, -1)
            dS = Pt * (dP - Drow[:, :, i0:i1].unsqueeze(-1)) * scale
            dqu[:, :, i0:i1] += dS @ k[:, :, j0:j1]
            dk[:, :, j0:j1] += dS.transpose(-2, -1) @ qu[:, :, i0:i1]
            dR = torch.zeros_like(R)
            dR.scatter_add_(-1, col.expand(B, H, i1 - i0, j1 - j0), dS)
            dqv[:, :, i0:i1] += dR @ band[None]
            contrib = (dR.transpose(-2, -1) @ qv[:, :, i0:i1]).sum(0)    # (H, nband, d)
            inside = (torch.arange(rbase, rbase + nband) >= 0) & (torch.arange(rbase, rbase + nband) <= 2 * T - 2)
            dp.index_add_(1, rows[inside], contrib[:, inside])
    return dqu + dqv, dk, dv, dp, dqu.sum((0, 2)), dqv.sum((0, 2))
