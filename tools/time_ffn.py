"""Fused-FFN kernel timing (graph replay, warm) and in-kernel phase stamps, tf32 and bf16 modes.
usage (GPU box): python tools/time_ffn.py [M]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tailored_avsr_b200 import _lib, ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
lib = _lib.load()


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for dt in (torch.float32, torch.bfloat16):
    xn, x = rn(M, 256).to(dt), rn(M, 256)
    w1, b1, w2, b2 = (rn(2048, 256) / 16).to(dt), rn(2048), (rn(256, 2048) / 45).to(dt), rn(256)
    gA = rn(256)
    o1 = torch.empty(M, 256, device="cuda")
    o2, o3 = (torch.empty(M, 256, device="cuda", dtype=dt) for _ in range(2))

    def ffn():
        ops.ffn_fused(xn, w1, b1, w2, b2, 1, residual=x, alpha=0.5, out_main=o1, lnA=(gA, gA), out_lnA=o2,
                      lnB=(gA, gA), out_lnB=o3)

    h = torch.empty(M, 2048, device="cuda", dtype=dt)

    def two():
        ops.gemm_bias_act(xn, w1, b1, act=1, out=h)
        ops.gemm_rowln(h, w2, b2, residual=x, alpha=0.5, out_main=o1, lnA=(gA, gA), out_lnA=o2, lnB=(gA, gA),
                       out_lnB=o3)

    print(f"[{dt}] M={M} fused us {t(ffn):.1f}   two-kernel us {t(two):.1f}")
    # ---- phase breakdown from in-kernel globaltimer stamps ----
    nblk = 2 * ((M + 127) // 128)
    dbg = torch.zeros(nblk * 8, dtype=torch.int64, device="cuda")
    lib.tavsr_debug_set_ptr(dbg.data_ptr())
    ffn(); torch.cuda.synchronize()
    dbg.zero_()
    ffn(); torch.cuda.synchronize()
    lib.tavsr_debug_set_ptr(None)
    d = dbg.cpu().numpy().reshape(nblk, 8).astype(np.float64)
    t0 = d[:, 0].min()
    names = ["setup done (after pdl_wait)", "first h_full", "d_full (main loop done)", "cluster sync 1",
             "cluster sync 2 (exchange)", "finish", "second h_full"]
    print(f"--- {nblk} CTAs; kernel span {(d[:, 5].max() - t0) / 1e3:.1f} us; start spread {(d[:, 0].max() - t0) / 1e3:.1f} us")
    for i, n in enumerate(names):
        col = d[:, i]
        print(f"   {n:30s} median +{(np.median(col) - t0) / 1e3:7.1f} us   max +{(col.max() - t0) / 1e3:7.1f} us")
