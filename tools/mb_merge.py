"""Where the merge GEMM's time goes: back-to-back timing of gemm_rowln variants at C2 shape."""
import os, sys
import torch
sys.path.insert(0, os.getcwd())
from tailored_avsr_b200 import ops
DEV = "cuda"
B, T = 32, 250
M = B * T
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(DEV)
reps = 200

def bench(name, fn):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us", flush=True)

ctx, u = rn(M, 256), rn(M, 1024)
res, gA = rn(M, 256), rn(256)
w1 = torch.rand(B, device=DEV)
o1, o2 = (torch.empty(M, 256, device=DEV) for _ in range(2))
sb = (rn(256), rn(256))
for K2 in (1024, 512, 256):
    w = rn(256, 256 + K2) / 30
    uu = u[:, :K2]
    bench(f"merge K=256+{K2} full (res, main, lnA)", lambda: ops.gemm_rowln(ctx, w, gA, x2=uu, k1=256, segbias=sb, rowscale=(w1, 1 - w1), rows_per_seg=T, residual=res, out_main=o1, lnA=(gA, gA), out_lnA=o2))
w = rn(256, 1280) / 30
bench("merge K=1280 no lnA", lambda: ops.gemm_rowln(ctx, w, gA, x2=u, k1=256, segbias=sb, rowscale=(w1, 1 - w1), rows_per_seg=T, residual=res, out_main=o1))
bench("merge K=1280 no residual, no lnA", lambda: ops.gemm_rowln(ctx, w, gA, x2=u, k1=256, segbias=sb, rowscale=(w1, 1 - w1), rows_per_seg=T, out_main=o1))
x5 = rn(M, 1280)
bench("plain rowln K=1280 (res, main, lnA)", lambda: ops.gemm_rowln(x5, w, gA, residual=res, out_main=o1, lnA=(gA, gA), out_lnA=o2))
bench("plain rowln K=1280 main only", lambda: ops.gemm_rowln(x5, w, gA, out_main=o1))
x2 = rn(M, 256); w2 = rn(256, 256) / 16
bench("plain rowln K=256 main only", lambda: ops.gemm_rowln(x2, w2, gA, out_main=o1))
bench("plain rowln K=256 (res, main, lnA)", lambda: ops.gemm_rowln(x2, w2, gA, residual=res, out_main=o1, lnA=(gA, gA), out_lnA=o2))
wt = rn(256, 1280) / 30
ot = torch.empty(M, 256, device=DEV)
bench("tiled gemm K=1280 N=256", lambda: ops.gemm_bias_act(x5, wt, gA, out=ot))
