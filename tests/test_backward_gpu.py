"""GPU tests of the backward building blocks (csrc/backward.cu: transpose, column sums, activation
/ LayerNorm / CSGU backward) against the autograd-verified formulas of oracle/bwd_formulas.py."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_transpose_and_col_sums():
    from tailored_avsr_b200 import ops_backward as ob
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 300, generator=g).to(DEV)
    assert torch.equal(ob.transpose_2d(x), x.t().contiguous())
    y = torch.randn(1000, 300, generator=g).to(DEV)
    assert _rel(ob.col_sums(x), x.double().sum(0)) < 1e-5
    assert _rel(ob.col_sums(x, y), (x.double() * y.double()).sum(0)) < 1e-5


@pytest.mark.parametrize("act", [1, 2, 3])
def test_act_bwd(act):
    from oracle import bwd_formulas as bw
    from tailored_avsr_b200 import ops_backward as ob
    g = torch.Generator().manual_seed(act)
    z = (torch.randn(777, 2048, generator=g) * 2).to(DEV)
    dh = torch.randn(777, 2048, generator=g).to(DEV)
    got = ob.act_bwd(z, dh, act)
    zd, dd = z.double().cpu(), dh.double().cpu()
    want = {1: bw.swish_bwd, 2: bw.gelu_bwd, 3: lambda a, b: b * (a > 0)}[act](zd, dd)
    assert _rel(got, want) < 1e-5


@pytest.mark.parametrize("M,D", [(1000, 256), (77, 1024), (8, 512)])
def test_layernorm_bwd(M, D):
    from oracle import bwd_formulas as bw
    from tailored_avsr_b200 import ops_backward as ob
    g = torch.Generator().manual_seed(M + D)
    x = (torch.randn(M, D, generator=g) * 1.5 + 0.3).to(DEV)
    gam = torch.randn(D, generator=g).to(DEV)
    dy = torch.randn(M, D, generator=g).to(DEV)
    dres = torch.randn(M, D, generator=g).to(DEV)
    dx, dg, db = ob.layernorm_bwd(x, gam, dy, eps=1e-12, dres=dres)
    wx, wg, wb = bw.layernorm_bwd(x.double().cpu(), gam.double().cpu(), dy.double().cpu())
    assert _rel(dx, wx + dres.double().cpu()) < 2e-5
    assert _rel(dg, wg) < 2e-5 and _rel(db, wb) < 2e-5


@pytest.mark.parametrize("B,T", [(2, 64), (3, 100), (1, 7)])
def test_csgu_bwd(B, T):
    from oracle import bwd_formulas as bw
    from tailored_avsr_b200 import ops, ops_backward as ob
    Ch = 256
    g = torch.Generator().manual_seed(B * T)
    h = torch.randn(B * T, 2 * Ch, generator=g).to(DEV)
    ng, nb = torch.randn(Ch, generator=g).to(DEV), torch.randn(Ch, generator=g).to(DEV)
    cw = (torch.randn(Ch, 31, generator=g) * 0.2).to(DEV)
    cb = torch.randn(Ch, generator=g).to(DEV)
    du = torch.randn(B * T, Ch, generator=g).to(DEV)
    stats = torch.empty(B * T, 2, device=DEV)
    ops.csgu(h, ng, nb, cw, cb, B, T, round_out=False, stats=stats)      # forward fills (mean, rstd)
    dh, dng, dnb, dcw, dcb = ob.csgu_bwd(h, ng, nb, cw, cb, stats, du, B, T)
    c = lambda t: t.double().cpu()                                        # noqa: E731
    want = bw.csgu_bwd(c(h).view(B, T, 2 * Ch), c(ng), c(nb), c(cw).view(Ch, 1, 31), c(cb),
                       c(du).view(B, T, Ch))
    assert _rel(dh, want[0].reshape(B * T, 2 * Ch)) < 5e-5
    assert _rel(dng, want[1]) < 5e-5 and _rel(dnb, want[2]) < 5e-5
    assert _rel(dcw, want[3].reshape(Ch, 31)) < 5e-5 and _rel(dcb, want[4]) < 5e-5


def test_merge_learned_ave_bwd():
    from oracle import bwd_formulas as bw
    from tailored_avsr_b200 import ops_backward as ob
    B, T, D = 4, 77, 256
    g = torch.Generator().manual_seed(7)
    lens = torch.tensor([77, 40, 1, 0], dtype=torch.int32)
    x1, x2, dm = (torch.randn(B * T, D, generator=g).to(DEV) for _ in range(3))
    a1, b1, a2, b2 = (torch.randn(D, generator=g).to(DEV) for _ in range(4))
    c1, e1, c2, e2 = 0.3, -0.2, 0.1, 0.7
    dx1, dx2, grads = ob.merge_learned_ave_bwd(x1, x2, dm, lens.to(DEV), a1, c1, b1, e1, a2, c2, b2, e2, B, T)
    c = lambda t: t.double().cpu()                                        # noqa: E731
    s_ = lambda v: torch.tensor(v, dtype=torch.float64)                   # noqa: E731
    outs = bw.learned_ave_merge_bwd(c(x1).view(B, T, D), c(x2).view(B, T, D), lens.long(),
                                    c(a1), s_(c1), c(b1), s_(e1), c(a2), s_(c2), c(b2), s_(e2),
                                    (c(dm).view(B, T, D),))
    (wx1, wa1, wc1, wb1, we1), (wx2, wa2, wc2, wb2, we2) = outs
    assert _rel(dx1, wx1.reshape(B * T, D)) < 5e-5 and _rel(dx2, wx2.reshape(B * T, D)) < 5e-5
    gr = grads.double().cpu()
    for got, want in ((gr[0:256], wa1), (gr[256:512], wb1), (gr[512:768], wa2), (gr[768:1024], wb2)):
        assert _rel(got, want) < 5e-5
    for got, want in ((gr[1024], wc1), (gr[1025], we1), (gr[1026], wc2), (gr[1027], we2)):
        assert abs(float(got) - float(want)) <= 5e-5 * max(1.0, abs(float(want)))


@pytest.mark.parametrize("B,T,lens", [(2, 64, [64, 40]), (3, 100, [100, 1, 77]), (2, 250, [250, 130]),
                                      (2, 130, [0, 130]), (1, 17, [17])])
def test_relpos_attention_bwd(B, T, lens):
    """tavsr_relpos_attn_bwd (P recomputed from the forward's log-sum-exp, fp32 FMA tiles, atomics for
    the d q parts and d pos) against the autograd-verified dense formula in fp64."""
    from oracle import bwd_formulas as bw
    from tailored_avsr_b200 import ops, ops_backward as ob
    H, dk = 4, 64
    g = torch.Generator().manual_seed(T + B)
    qkv = torch.randn(B * T, 3 * H * dk, generator=g)
    pos = torch.randn(2 * T - 1, H * dk, generator=g)
    u = torch.randn(H * dk, generator=g) * 0.5
    v = torch.randn(H * dk, generator=g) * 0.5
    dctx = torch.randn(B * T, H * dk, generator=g)
    lens_t = torch.tensor(lens, dtype=torch.int32)
    lse = torch.empty(B, H, T, device=DEV)
    ctx = ops.relpos_attn(qkv.to(DEV), pos.to(DEV), u.to(DEV), v.to(DEV), lens_t.to(DEV), B, T, H,
                          round_out=False, lse=lse)
    dqkv, dpos, du, dv = ob.relpos_attn_bwd(qkv.to(DEV), pos.to(DEV), u.to(DEV), v.to(DEV),
                                            lens_t.to(DEV), ctx, dctx.to(DEV), lse, B, T, H)
    torch.cuda.synchronize()
    split = lambda t: t.double().view(B, T, H, dk).transpose(1, 2)            # noqa: E731
    q, k, vv = [split(t) for t in qkv.split(H * dk, dim=1)]
    p = pos.double().view(2 * T - 1, H, dk).transpose(0, 1)
    wq, wk, wv, wp, wu, wvb = bw.relpos_attn_core_bwd(q, k, vv, p, u.double().view(H, dk),
                                                      v.double().view(H, dk), lens_t.long(), split(dctx))
    unsplit = lambda t: t.transpose(1, 2).reshape(B * T, H * dk)             # noqa: E731
    gq, gk, gv = dqkv.split(H * dk, dim=1)
    tol = 5e-3   # the forward's scores / lse are TF32, the recomputation is fp32
    assert _rel(gv, unsplit(wv)) < tol, _rel(gv, unsplit(wv))
    assert _rel(gk, unsplit(wk)) < tol, _rel(gk, unsplit(wk))
    assert _rel(gq, unsplit(wq)) < tol, _rel(gq, unsplit(wq))
    assert _rel(dpos, wp.transpose(0, 1).reshape(2 * T - 1, H * dk)) < tol
    assert _rel(du, wu.reshape(-1)) < tol and _rel(dv, wvb.reshape(-1)) < tol


@pytest.mark.parametrize("M,N,K", [(8000, 2048, 256), (385, 256, 1024), (1000, 768, 256)])
def test_linear_bwd_on_the_tcgen05_gemm(M, N, K):
    from tailored_avsr_b200 import ops_backward as ob
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    dy = torch.randn(M, N, generator=g).to(DEV)
    dx, dw, db = ob.linear_bwd(x, w, dy)
    assert _rel(dx, dy.double() @ w.double()) < 2e-3
    assert _rel(dw, dy.double().t() @ x.double()) < 2e-3
    assert _rel(db, dy.double().sum(0)) < 1e-5


@pytest.mark.parametrize("act", [1, 2, 3])
def test_act_fwd_with_and_without_mask(act):
    from tailored_avsr_b200 import ops_backward as ob
    g = torch.Generator().manual_seed(act)
    z = torch.randn(333, 512, generator=g).to(DEV) * 2
    mask = (torch.rand(333, 512, generator=g) > 0.1).float().to(DEV) / 0.9
    zd = z.double()
    want = {1: zd * torch.sigmoid(zd), 2: torch.nn.functional.gelu(zd), 3: torch.relu(zd)}[act]
    assert _rel(ob.act_fwd(z, act), want) < 1e-6
    assert _rel(ob.act_fwd(z, act, mask=mask), want * mask.double()) < 1e-6
