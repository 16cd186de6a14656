"""Run each hot kernel a few times at the C2 shapes (for ncu captures).
usage: python tools/prof_ops.py [gemm_w1|gemm_rowln|csgu|attn|ctc|merge|all]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tailored_avsr_b200 import ops  # noqa: E402

DEV = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, T = 32, 250
M = B * T
g = torch.Generator().manual_seed(0)


def rn(*s):
    return torch.randn(*s, generator=g).to(DEV)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def run(fn):
    for _ in range(reps):
        flush.zero_()
        fn()
    torch.cuda.synchronize()


if which in ("gemm_w1", "all"):
    x, w, b = rn(M, 256), rn(2048, 256) / 16, rn(2048)
    out = torch.empty(M, 2048, device=DEV)
    run(lambda: ops.gemm_bias_act(x, w, b, act=1, out=out))
if which in ("gemm_gelu", "all"):
    x, w, b = rn(M, 256), rn(2048, 256) / 16, rn(2048)
    out = torch.empty(M, 2048, device=DEV)
    run(lambda: ops.gemm_bias_act(x, w, b, act=2, out=out))
if which in ("gemm_qkv", "all"):
    x, w, b = rn(M, 256), rn(768, 256) / 16, rn(768)
    out = torch.empty(M, 768, device=DEV)
    run(lambda: ops.gemm_bias_act(x, w, b, act=0, out=out))
if which in ("gemm_rowln", "all"):
    x, w, b = rn(M, 2048), rn(256, 2048) / 45, rn(256)
    res, gA = rn(M, 256), rn(256)
    o1, o2, o3 = (torch.empty(M, 256, device=DEV) for _ in range(3))
    run(lambda: ops.gemm_rowln(x, w, b, residual=res, alpha=0.5, out_main=o1, lnA=(gA, gA), out_lnA=o2,
                               lnB=(gA, gA), out_lnB=o3))
if which in ("gemm_merge", "all"):
    x1, x2, w, b = rn(M, 256), rn(M, 256), rn(256, 256) / 16, rn(256)
    res, gA = rn(M, 256), rn(256)
    w1 = torch.rand(B, device=DEV)
    o1, o2 = (torch.empty(M, 256, device=DEV) for _ in range(2))
    run(lambda: ops.gemm_rowln(x1, w, b, x2=x2, rowscale=(w1, 1 - w1), rows_per_seg=T, residual=res,
                               out_main=o1, lnA=(gA, gA), out_lnA=o2))
if which in ("ffn", "all"):
    xn, x = rn(M, 256), rn(M, 256)
    w1, b1, w2, b2 = rn(2048, 256) / 16, rn(2048), rn(256, 2048) / 45, rn(256)
    gA = rn(256)
    o1, o2, o3 = (torch.empty(M, 256, device=DEV) for _ in range(3))
    run(lambda: ops.ffn_fused(xn, w1, b1, w2, b2, 1, residual=x, alpha=0.5, out_main=o1, lnA=(gA, gA),
                              out_lnA=o2, lnB=(gA, gA), out_lnB=o3))
if which in ("csgu", "all"):
    h, ng = rn(M, 2048), rn(1024)
    cw = rn(1024, 31)
    o, st = torch.empty(M, 1024, device=DEV), torch.empty(M, 2, device=DEV)
    run(lambda: ops.csgu(h, ng, ng, cw, ng, B, T, out=o, stats=st))
if which in ("attn", "all"):
    qkv, pos, u = rn(M, 768), rn(2 * T - 1, 256), rn(256)
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
    o = torch.empty(M, 256, device=DEV)
    run(lambda: ops.relpos_attn(qkv, pos, u, u, lens, B, T, 4, out=o))
if which in ("ctc", "all"):
    logp = torch.log_softmax(rn(B, T, 41), -1)
    tg = torch.randint(1, 41, (B, 100), device=DEV)
    tl = torch.full((B,), 100, dtype=torch.int32, device=DEV)
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
    run(lambda: ops.ctc_loss(logp, tg, lens, tl))
    hs, w, bb = rn(M, 256), rn(41, 256), rn(41)
    run(lambda: ops.ctc_head(hs, w, bb, True, False, True))
if which in ("merge", "all"):
    d1, d2 = rn(M, 2), rn(M, 2)
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
    run(lambda: ops.merge_weights(d1, d2, lens, 0.1, 0.1, 0.1, 0.1, 256, B, T))
print("done", which)
