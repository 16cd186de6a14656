"""Warm back-to-back timing of the hot kernels at the C2 shapes (CUDA events around `reps`
launches of one op, PDL on): python tools/microbench.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
from tailored_avsr_b200 import ops  # noqa: E402

DEV = "cuda"
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B, T = 32, 250
M = B * T
g = torch.Generator().manual_seed(0)


def rn(*s):
    return torch.randn(*s, generator=g).to(DEV)


def bench(name, fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us", flush=True)


x, w, b = rn(M, 256), rn(2048, 256) / 16, rn(2048)
out = torch.empty(M, 2048, device=DEV)
bench("gemm_gelu 256->2048", lambda: ops.gemm_bias_act(x, w, b, act=2, out=out))
if hasattr(ops, "gemm_bias_act_stats"):
    bench("gemm_gelu_stats", lambda: ops.gemm_bias_act_stats(x, w, b, 2, 1024, out=out))
wq, bq = rn(768, 256) / 16, rn(768)
oq = torch.empty(M, 768, device=DEV)
bench("gemm_qkv 256->768", lambda: ops.gemm_bias_act(x, wq, bq, act=0, out=oq))
h, ng, cw = rn(M, 2048), rn(1024), rn(1024, 31)
o, st = torch.empty(M, 1024, device=DEV), torch.empty(M, 2, device=DEV)
bench("csgu (stats+conv)", lambda: ops.csgu(h, ng, ng, cw, ng, B, T, out=o, stats=st))
if hasattr(ops, "csgu_fused"):
    part = torch.rand(M, 8, 2, device=DEV)
    bench("csgu_fused no dots", lambda: ops.csgu_fused(h, ng, ng, cw, ng, B, T, part, 8, 128))
    bench("csgu_fused dots", lambda: ops.csgu_fused(h, ng, ng, cw, ng, B, T, part, 8, 128, dots=(ng, ng)))
qkv, pos, u = rn(M, 768), rn(2 * T - 1, 256), rn(256)
lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
oc = torch.empty(M, 256, device=DEV)
bench("attn", lambda: ops.relpos_attn(qkv, pos, u, u, lens, B, T, 4, out=oc))
if hasattr(ops, "merge_weights2"):
    bench("attn dots", lambda: ops.relpos_attn(qkv, pos, u, u, lens, B, T, 4, out=oc, dots=(u, u)))
ctx, uu, va, vb = rn(M, 256), rn(M, 1024), rn(256), rn(1024)
bench("row_dots", lambda: ops.row_dots(ctx, va, va, uu, vb, vb))
d1, d2 = rn(M, 2), rn(M, 2)
bench("merge_weights", lambda: ops.merge_weights(d1, d2, lens, .1, .2, .3, .4, 256, B, T))
if hasattr(ops, "merge_weights2"):
    p1, p2 = rn(M, 8, 2), rn(M, 8, 2)
    bench("merge_weights2 np=8", lambda: ops.merge_weights2(p1, 8, p2, 8, lens, None, .1, .2, .3, .4, 256, B, T))
