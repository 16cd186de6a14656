"""Per-kernel checks of the bf16 mode (kind::f16 MMAs on bf16 operands, fp32 accumulation, bf16
storage of every tensor that only feeds the next tensor-core product) and of the tf32x3 split.

Each kernel is compared with an fp64 evaluation of the SAME bf16-rounded inputs, so the bound only
has to cover fp32 accumulation order and the bf16 rounding of the stored outputs (2^-9 relative per
element, ~2.3e-3 Frobenius on random data) - not the operand rounding the mode accepts by design.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def _ops():
    from tailored_avsr_b200 import ops
    return ops


def rel_fro(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _ln(x, g, b, eps):
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


OUT_BF16_FRO = 3e-3   # a bf16-stored output: rounding 2^-9 per element
OUT_F32_FRO = 5e-5    # an fp32 output of exact bf16 products: fp32 summation order only


@pytest.mark.parametrize("M,N,K", [(8000, 2048, 256), (8000, 768, 256), (300, 256, 1024), (77, 48, 64),
                                   (1992, 1024, 2304), (130, 96, 256), (499, 3072, 256)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_bf16_out(M, N, K, act):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K + act)
    x = torch.randn(M, K, generator=g).to(DEV).to(BF)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).to(BF)
    b = torch.randn(N, generator=g).to(DEV)
    y = ops.gemm_bias_act(x, w, b, act=act, out_dtype=BF)
    assert y.dtype == BF and y.shape == (M, N)
    ref = x.double() @ w.double().t() + b.double()
    if act == 1:
        ref = ref * torch.sigmoid(ref)
    elif act == 2:
        ref = F.gelu(ref)
    assert rel_fro(y, ref) < OUT_BF16_FRO, rel_fro(y, ref)
    assert max_rel(y, ref) < 8e-3


@pytest.mark.parametrize("M,K", [(128, 256), (1992, 2048), (300, 1024), (8000, 512), (64, 4864)])
@pytest.mark.parametrize("outs", ["f32", "bf16", "mixed"])
def test_rowln_bf16(M, K, outs):
    """Row-complete GEMM on bf16 operands: main (residual stream) fp32 or bf16, LayerNorm outputs
    fp32 or bf16, chained LayerNorm."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g).to(DEV).to(BF)
    w = (torch.randn(256, K, generator=g) / math.sqrt(K)).to(DEV).to(BF)
    b = torch.randn(256, generator=g).to(DEV)
    res = torch.randn(M, 256, generator=g).to(DEV)
    g0, b0, gA, bA, gB, bB = [torch.randn(256, generator=g).to(DEV) for _ in range(6)]
    dm = BF if outs == "bf16" else torch.float32
    dA = torch.float32 if outs == "f32" else BF
    dB = BF if outs == "bf16" else torch.float32
    main = torch.empty(M, 256, device=DEV, dtype=dm)
    oA = torch.empty(M, 256, device=DEV, dtype=dA)
    oB = torch.empty(M, 256, device=DEV, dtype=dB)
    ops.gemm_rowln(x, w, b, residual=res, alpha=0.5, out_main=main, lnA=(gA, bA), out_lnA=oA,
                   lnB=(gB, bB), out_lnB=oB, eps=1e-12)
    v = res.double() + 0.5 * (x.double() @ w.double().t() + b.double())
    tol = lambda d: OUT_BF16_FRO if d == BF else OUT_F32_FRO  # noqa: E731
    assert rel_fro(main, v) < tol(dm), rel_fro(main, v)
    assert rel_fro(oA, _ln(v, gA.double(), bA.double(), 1e-12)) < tol(dA) + 2e-5
    assert rel_fro(oB, _ln(v, gB.double(), bB.double(), 1e-12)) < tol(dB) + 2e-5
    # chained: norm_final then the next block's LayerNorm
    ops.gemm_rowln(x, w, b, residual=res, alpha=0.5, ln0=(g0, b0), eps0=1e-12, out_main=main,
                   lnA=(gA, bA), out_lnA=oA)
    v1 = _ln(v, g0.double(), b0.double(), 1e-12)
    assert rel_fro(main, v1) < tol(dm) + 2e-5
    assert rel_fro(oA, _ln(v1, gA.double(), bA.double(), 1e-12)) < tol(dA) + 5e-5


@pytest.mark.parametrize("B,T,K1,K2", [(5, 77, 256, 1024), (3, 250, 256, 1024), (2, 40, 64, 128),
                                       (4, 61, 256, 256)])
def test_rowln_bf16_sequential_dual(B, T, K1, K2):
    ops = _ops()
    g = torch.Generator().manual_seed(B * T + K2)
    M = B * T
    x1 = torch.randn(M, K1, generator=g).to(DEV).to(BF)
    x2 = torch.randn(M, K2, generator=g).to(DEV).to(BF)
    w = torch.cat([torch.randn(256, K1, generator=g) / math.sqrt(K1),
                   torch.randn(256, K2, generator=g) / math.sqrt(K2)], 1).contiguous().to(DEV).to(BF)
    b, c1, c2 = [torch.randn(256, generator=g).to(DEV) for _ in range(3)]
    res = torch.randn(M, 256, generator=g).to(DEV)
    w1 = torch.rand(B, generator=g).to(DEV)
    w2 = 1 - w1
    gA, bA = torch.randn(256, generator=g).to(DEV), torch.randn(256, generator=g).to(DEV)
    main = torch.empty(M, 256, device=DEV)
    oA = torch.empty(M, 256, device=DEV, dtype=BF)
    ops.gemm_rowln(x1, w, b, x2=x2, k1=K1, segbias=(c1, c2), rowscale=(w1, w2), rows_per_seg=T,
                   residual=res, alpha=1.0, out_main=main, lnA=(gA, bA), out_lnA=oA)
    s1 = w1.double().repeat_interleave(T)[:, None]
    s2 = w2.double().repeat_interleave(T)[:, None]
    wd = w.double()
    v = (res.double() + s1 * (x1.double() @ wd[:, :K1].t() + c1.double())
         + s2 * (x2.double() @ wd[:, K1:].t() + c2.double()) + b.double())
    assert rel_fro(main, v) < OUT_F32_FRO, rel_fro(main, v)
    assert rel_fro(oA, _ln(v, gA.double(), bA.double(), 1e-12)) < OUT_BF16_FRO


@pytest.mark.parametrize("M", [128, 8000, 1992, 77, 333])
@pytest.mark.parametrize("variant", ["macaron", "final", "nores"])
def test_ffn_fused_bf16(M, variant):
    """Fused FFN in bf16 mode: bf16 xn / W1 / W2, the Swish'd hidden packed as bf16 pairs in TMEM
    (A operand of the second GEMM).  The reference rounds the hidden to bf16 like the kernel does."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + len(variant))
    xn = torch.randn(M, 256, generator=g).to(DEV).to(BF)
    x = torch.randn(M, 256, generator=g).to(DEV)
    w1 = (torch.randn(2048, 256, generator=g) / 16).to(DEV).to(BF)
    b1 = torch.randn(2048, generator=g).to(DEV) * 0.1
    w2 = (torch.randn(256, 2048, generator=g) / 45).to(DEV).to(BF)
    b2 = torch.randn(256, generator=g).to(DEV) * 0.1
    g0, b0, gA, bA, gB, bB = [torch.randn(256, generator=g).to(DEV) for _ in range(6)]
    h = (xn.double() @ w1.double().t() + b1.double())
    h = (h * torch.sigmoid(h)).float().to(BF).double()      # the packed hidden is bf16
    ffn = h @ w2.double().t() + b2.double()
    main = torch.empty(M, 256, device=DEV)
    oA = torch.empty(M, 256, device=DEV, dtype=BF)
    oB = torch.empty(M, 256, device=DEV, dtype=BF)
    # hidden values at a bf16 rounding boundary may round the other way (fp32 vs fp64 pre-activation)
    tol_main = 3e-4
    if variant == "macaron":
        v0 = x.double() + 0.5 * ffn
        ops.ffn_fused(xn, w1, b1, w2, b2, 1, residual=x, alpha=0.5, out_main=main, lnA=(gA, bA),
                      out_lnA=oA, lnB=(gB, bB), out_lnB=oB)
        assert rel_fro(main, v0) < tol_main, rel_fro(main, v0)
        assert rel_fro(oA, _ln(v0, gA.double(), bA.double(), 1e-12)) < OUT_BF16_FRO
        assert rel_fro(oB, _ln(v0, gB.double(), bB.double(), 1e-12)) < OUT_BF16_FRO
    elif variant == "final":
        v0 = x.double() + 0.5 * ffn
        oAf = torch.empty(M, 256, device=DEV)        # after_norm: fp32 encoder output
        ops.ffn_fused(xn, w1, b1, w2, b2, 1, residual=x, alpha=0.5, ln0=(g0, b0), out_main=main,
                      lnA=(gA, bA), out_lnA=oAf)
        v1 = _ln(v0, g0.double(), b0.double(), 1e-12)
        assert rel_fro(main, v1) < 2 * tol_main, rel_fro(main, v1)
        assert rel_fro(oAf, _ln(v1, gA.double(), bA.double(), 1e-12)) < 3 * tol_main
    else:   # the fusion module's form: no residual, alpha 1, LN0 only
        ops.ffn_fused(xn, w1, b1, w2, b2, 1, residual=None, alpha=1.0, ln0=(g0, b0), out_main=main)
        assert rel_fro(main, _ln(ffn, g0.double(), b0.double(), 1e-12)) < 2 * tol_main


def _rel_shift(x):
    b, h, t, n = x.shape
    zero_pad = torch.zeros((b, h, t, 1), dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=-1).view(b, h, n + 1, t)
    return x_padded[:, :, 1:].view_as(x)[:, :, :, : n // 2 + 1]


@pytest.mark.parametrize("B,T,lens", [(2, 64, [64, 40]), (3, 100, [100, 1, 77]), (2, 250, [250, 130]),
                                      (1, 17, [17]), (2, 130, [0, 130]), (1, 300, [300]),
                                      (2, 515, [515, 260]), (1, 1500, [1333])])
def test_relpos_attention_bf16(B, T, lens):
    ops = _ops()
    H, dk = 4, 64
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B * T, 3 * H * dk, generator=g).to(BF)
    pos = torch.randn(2 * T - 1, H * dk, generator=g).to(BF)
    u = torch.randn(H * dk, generator=g) * 0.5
    v = torch.randn(H * dk, generator=g) * 0.5
    lens_t = torch.tensor(lens, dtype=torch.int32)
    out = ops.relpos_attn(qkv.to(DEV), pos.to(DEV), u.to(DEV), v.to(DEV), lens_t.to(DEV), B, T, H)
    assert out.dtype == BF
    # espnet-style reference in fp64 on the bf16 values
    q, k, vv = [t.double().view(B, T, H, dk).transpose(1, 2) for t in qkv.split(H * dk, dim=1)]
    p = pos.double().view(1, 2 * T - 1, H, dk).transpose(1, 2)
    ac = (q + u.double().view(1, H, 1, dk)) @ k.transpose(-2, -1)
    bd = _rel_shift((q + v.double().view(1, H, 1, dk)) @ p.transpose(-2, -1))
    scores = (ac + bd) / math.sqrt(dk)
    mask = (torch.arange(T)[None, :] >= lens_t[:, None].long())[:, None, None, :]
    scores = scores.masked_fill(mask, torch.finfo(torch.float64).min)
    attn = torch.softmax(scores, dim=-1).masked_fill(mask, 0.0)
    ref = (attn @ vv).transpose(1, 2).reshape(B * T, H * dk)
    # probabilities are rounded to bf16 before P.V (2^-9 each, averaged over the keys) and the
    # context is stored as bf16
    assert rel_fro(out, ref) < 6e-3, rel_fro(out, ref)
    assert max_rel(out, ref) < 2e-2


@pytest.mark.parametrize("B,T", [(2, 64), (3, 100), (2, 250), (1, 7)])
def test_csgu_bf16(B, T):
    ops = _ops()
    Ch = 1024
    g = torch.Generator().manual_seed(B * T)
    h = torch.randn(B * T, 2 * Ch, generator=g).to(BF)
    ng, nb = torch.randn(Ch, generator=g), torch.randn(Ch, generator=g)
    cw = torch.randn(Ch, 1, 31, generator=g) * 0.2
    cb = torch.randn(Ch, generator=g)
    out = ops.csgu(h.to(DEV), ng.to(DEV), nb.to(DEV), cw.reshape(Ch, 31).to(DEV).contiguous(),
                   cb.to(DEV), B, T)
    assert out.dtype == BF
    hd = h.double().view(B, T, 2 * Ch)
    r, gt = hd.chunk(2, dim=-1)
    gt = F.layer_norm(gt, (Ch,), ng.double(), nb.double(), 1e-12)
    gt = F.conv1d(gt.transpose(1, 2), cw.double(), cb.double(), padding=15, groups=Ch).transpose(1, 2)
    ref = (r * gt).reshape(B * T, Ch)
    assert rel_fro(out, ref) < OUT_BF16_FRO, rel_fro(out, ref)
    assert float((out.double().cpu() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max()) + 1e-6


@pytest.mark.parametrize("M", [1, 333, 8000])
def test_row_dots_bf16(M):
    ops = _ops()
    g = torch.Generator().manual_seed(M)
    a1 = torch.randn(M, 256, generator=g).to(DEV).to(BF)
    a2 = torch.randn(M, 1024, generator=g).to(DEV).to(BF)
    v = [torch.randn(k, generator=g).to(DEV) for k in (256, 256, 1024, 1024)]
    o1, o2 = ops.row_dots(a1, v[0], v[1], a2, v[2], v[3])
    r1 = torch.stack([a1.double() @ v[0].double(), a1.double() @ v[1].double()], 1)
    r2 = torch.stack([a2.double() @ v[2].double(), a2.double() @ v[3].double()], 1)
    assert (o1.cpu().double() - r1.cpu()).abs().max() < 1e-4 * 16
    assert (o2.cpu().double() - r2.cpu()).abs().max() < 1e-4 * 32


@pytest.mark.parametrize("M,D", [(1000, 256), (37, 1024)])
def test_layernorm_and_cast_bf16_outputs(M, D):
    ops = _ops()
    g = torch.Generator().manual_seed(M)
    x = (torch.randn(M, D, generator=g) * 3 + 1).to(DEV)
    gA, bA, gB, bB = [torch.randn(D, generator=g).to(DEV) for _ in range(4)]
    oB = torch.empty(M, D, device=DEV)
    oA = ops.layernorm(x, gA, bA, eps=1e-12, gB=gB, bB=bB, outB=oB, out_dtype=BF)
    wantA = _ln(x.double(), gA.double(), bA.double(), 1e-12)
    assert oA.dtype == BF and rel_fro(oA, wantA) < OUT_BF16_FRO
    assert max_rel(oB, _ln(x.double(), gB.double(), bB.double(), 1e-12)) < 1e-5
    c = ops.cast_bf16(x)
    assert c.dtype == BF and torch.equal(c, x.to(BF))


def test_vocab_residual_bf16_ln_and_conv2d_bf16_operand():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    M, D, V = 70, 256, 41
    x = torch.randn(M, D, generator=g)
    p = torch.randn(M, V, generator=g).softmax(-1)
    w = torch.randn(D, V, generator=g)
    b = torch.randn(D, generator=g)
    gam, bet = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    out, xn = ops.vocab_residual(x.to(DEV), p.to(DEV), w.to(DEV), b.to(DEV),
                                 ln=(gam.to(DEV), bet.to(DEV)), eps=1e-12, ln_dtype=BF)
    want = x + p @ w.t() + b
    assert (out.cpu() - want).abs().max() < 1e-5 * max(1.0, float(want.abs().max()))
    assert xn.dtype == BF and rel_fro(xn, F.layer_norm(want, (D,), gam, bet, 1e-12)) < OUT_BF16_FRO
    # conv2d front end: im2col operand written as bf16
    C, B, Tin, Fin = 256, 2, 67, 80
    xx = torch.randn(B, Tin, Fin, generator=g)
    w1 = torch.randn(C, 9, generator=g) / 3
    b1 = torch.randn(C, generator=g) * 0.1
    a32 = ops.conv2d_sub_im2col(xx.to(DEV), w1.to(DEV), b1.to(DEV))
    a16 = ops.conv2d_sub_im2col(xx.to(DEV), w1.to(DEV), b1.to(DEV), out_dtype=BF)
    assert a16.dtype == BF and torch.equal(a16, a32.to(BF))


@pytest.mark.parametrize("M,N,K", [(300, 256, 1024), (1000, 2048, 256), (77, 48, 64)])
def test_tf32x3_split_product_is_fp32_class(M, N, K):
    """[hi | hi | lo] . [hi | lo | hi]^T over the tripled reduction axis reproduces the fp32 product
    to ~1e-6 where a single TF32 pass gives ~5e-4."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    x3, w3 = ops.split_tf32(x, "x"), ops.split_tf32(w, "w")
    assert x3.shape == (M, 3 * K) and w3.shape == (N, 3 * K)
    assert torch.equal(x3[:, :K], x3[:, K:2 * K]) and torch.equal(x3[:, :K] + x3[:, 2 * K:], x)
    assert torch.equal(w3[:, :K], w3[:, 2 * K:]) and torch.equal(w3[:, :K] + w3[:, K:2 * K], w)
    y3 = ops.gemm_bias_act(x3, w3, b)
    y1 = ops.gemm_bias_act(x, w, b)
    ref = x.double() @ w.double().t() + b.double()
    e3, e1 = rel_fro(y3, ref), rel_fro(y1, ref)
    assert e3 < 2e-5 and e3 < e1 / 20, (e3, e1)
    if N == 256:
        res = torch.randn(M, 256, generator=g).to(DEV)
        main = torch.empty(M, 256, device=DEV)
        ops.gemm_rowln(x3, w3, b, residual=res, alpha=0.5, out_main=main)
        assert rel_fro(main, res.double() + 0.5 * ref) < 2e-5


@pytest.mark.parametrize("dt", [torch.float32, BF])
@pytest.mark.parametrize("B,T,lens", [(6, 90, [90, 1, 45, 0, 89, 33]), (3, 250, [250, 130, 61]),
                                      (2, 1500, [1500, 777]), (1, 5, [5])])
def test_merge_scores_cluster_kernel_equals_row_dots_plus_merge_weights(dt, B, T, lens):
    """tavsr_merge_scores (row dots + masked softmax pooling + 2-way softmax, a 4-CTA cluster per
    utterance with a DSMEM combine) against the two-kernel sequence and the fp64 formula."""
    ops = _ops()
    g = torch.Generator().manual_seed(B * T)
    a1 = (torch.randn(B * T, 256, generator=g) * 0.5).to(DEV).to(dt)
    a2 = (torch.randn(B * T, 1024, generator=g) * 0.5).to(DEV).to(dt)
    v = [(torch.randn(k, generator=g) * s).to(DEV) for k, s in ((256, 0.3), (256, 0.1), (1024, 0.2), (1024, 0.05))]
    lens_t = torch.tensor(lens, dtype=torch.int32).to(DEV)
    pb1, pb2, wb1, wb2 = 0.3, -0.2, 0.1, 0.7
    w1, w2 = ops.merge_scores(a1, a2, v[0], v[1], v[2], v[3], lens_t, pb1, pb2, wb1, wb2, 256, B, T)
    d1, d2 = ops.row_dots(a1, v[0], v[1], a2, v[2], v[3])
    r1, r2 = ops.merge_weights(d1, d2, lens_t, pb1, pb2, wb1, wb2, 256, B, T)
    assert float((w1 - r1).abs().max()) < 2e-5 and float((w2 - r2).abs().max()) < 2e-5
    om = []
    for a, va, vb, pb, wb in ((a1, v[0], v[1], pb1, wb1), (a2, v[2], v[3], pb2, wb2)):
        ad = a.double().cpu().view(B, T, -1)
        sc = (ad @ va.double().cpu() + pb) / 16.0
        mask = torch.arange(T)[None, :] >= torch.tensor(lens)[:, None]
        sc = sc.masked_fill(mask, torch.finfo(torch.float32).min)
        s = torch.softmax(sc, dim=-1).masked_fill(mask, 0.0)
        om.append((s * (ad @ vb.double().cpu())).sum(-1) + wb)
    ref = torch.softmax(torch.stack(om, dim=-1), dim=-1)
    assert float((w1.double().cpu() - ref[:, 0]).abs().max()) < 2e-5
    assert float((w2.double().cpu() - ref[:, 1]).abs().max()) < 2e-5


@pytest.mark.parametrize("dt", [torch.float32, BF])
@pytest.mark.parametrize("M", [8000, 300, 1992])
def test_gemm_group2_equals_the_two_projections(dt, M):
    """The grouped launch (QKV projection without activation + channel_proj1 with GELU over one tile
    schedule) returns exactly what the two stand-alone GEMMs return."""
    ops = _ops()
    g = torch.Generator().manual_seed(M)
    K, N1, N2 = 256, 768, 2048
    x1 = torch.randn(M, K, generator=g).to(DEV).to(dt)
    x2 = torch.randn(M, K, generator=g).to(DEV).to(dt)
    w1 = (torch.randn(N1, K, generator=g) / 16).to(DEV).to(dt)
    w2 = (torch.randn(N2, K, generator=g) / 16).to(DEV).to(dt)
    b1, b2 = torch.randn(N1, generator=g).to(DEV), torch.randn(N2, generator=g).to(DEV)
    y1, y2 = ops.gemm_group2(x1, w1, b1, x2, w2, b2, out_dtype=dt)
    r1 = ops.gemm_bias_act(x1, w1, b1, act=ops.ACT_NONE, out_dtype=dt)
    r2 = ops.gemm_bias_act(x2, w2, b2, act=ops.ACT_GELU, out_dtype=dt)
    assert y1.dtype == dt and y2.dtype == dt
    assert torch.equal(y1, r1) and torch.equal(y2, r2)
    ref2 = F.gelu(x2.double() @ w2.double().t() + b2.double())
    assert rel_fro(y2, ref2) < (OUT_BF16_FRO if dt == BF else 2e-3)


@pytest.mark.parametrize("dt", [torch.float32, BF])
@pytest.mark.parametrize("B,T", [(2, 64), (3, 100), (2, 250), (1, 7)])
def test_csgu_onepass_cluster_kernel_matches_the_two_kernel_sequence(dt, B, T):
    """Debug knob 10: LayerNorm statistics from the convolution's own tile (8-CTA cluster, DSMEM
    partials, Chan combine) - same output and the same (mean, rstd) as the two-kernel sequence."""
    from tailored_avsr_b200 import _lib
    ops = _ops()
    Ch = 1024
    g = torch.Generator().manual_seed(B * T)
    h = torch.randn(B * T, 2 * Ch, generator=g).to(DEV).to(dt)
    ng, nb = torch.randn(Ch, generator=g).to(DEV), torch.randn(Ch, generator=g).to(DEV)
    cw = (torch.randn(Ch, 31, generator=g) * 0.2).to(DEV)
    cb = torch.randn(Ch, generator=g).to(DEV)
    st_a = torch.empty(B * T, 2, device=DEV)
    st_b = torch.empty(B * T, 2, device=DEV)
    want = ops.csgu(h, ng, nb, cw, cb, B, T, round_out=False, stats=st_a)
    lib = _lib.load()
    lib.tavsr_debug_set(10, 1)
    try:
        got = ops.csgu(h, ng, nb, cw, cb, B, T, round_out=False, stats=st_b)
    finally:
        lib.tavsr_debug_set(10, 0)
    assert rel_fro(got, want) < (2e-3 if dt == BF else 1e-5), rel_fro(got, want)
    assert max_rel(st_b, st_a) < 1e-5
