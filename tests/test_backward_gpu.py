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
