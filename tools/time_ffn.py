import sys, torch
sys.path.insert(0, ".")
from tailored_avsr_b200 import ops, _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda()
xn, x = rn(M, 256), rn(M, 256)
w1, b1, w2, b2 = rn(2048, 256) / 16, rn(2048), rn(256, 2048) / 45, rn(256)
gA = rn(256)
o1, o2, o3 = (torch.empty(M, 256, device="cuda") for _ in range(3))


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


def ffn():
    ops.ffn_fused(xn, w1, b1, w2, b2, 1, residual=x, alpha=0.5, out_main=o1, lnA=(gA, gA), out_lnA=o2,
                  lnB=(gA, gA), out_lnB=o3)


lib = _lib.load()
lib.tavsr_debug_set(5, 2)
print(f"M={M} ffn pair (v2) us", t(ffn))
lib.tavsr_debug_set(5, 0)
print(f"M={M} ffn v1 us", t(ffn))
h = torch.empty(M, 2048, device="cuda")


def two():
    ops.gemm_bias_act(xn, w1, b1, act=1, out=h)
    ops.gemm_rowln(h, w2, b2, residual=x, alpha=0.5, out_main=o1, lnA=(gA, gA), out_lnA=o2, lnB=(gA, gA),
                   out_lnB=o3)


print(f"M={M} two-kernel us", t(two))

# ---- phase breakdown from in-kernel globaltimer stamps ----
import numpy as np
for ver in (2, 0):
    lib.tavsr_debug_set(5, ver)
    nblk = (4 * ((M + 255) // 256)) if ver == 2 else 2 * ((M + 127) // 128)
    dbg = torch.zeros(nblk * 8, dtype=torch.int64, device="cuda")
    lib.tavsr_debug_set_ptr(dbg.data_ptr())
    if ver == 2:
        lib.tavsr_debug_set(6, 1)
    ffn(); torch.cuda.synchronize()
    dbg.zero_()
    ffn(); torch.cuda.synchronize()
    lib.tavsr_debug_set_ptr(None)
    d = dbg.cpu().numpy().reshape(nblk, 8).astype(np.float64)
    t0 = d[:, 0].min()
    names = ["setup done", "first h_full", "d_full (main loop done)", "cluster sync 1", "cluster sync 2 (exchange)", "finish", "second h_full"]
    print(f"--- version {'v2 pair' if ver == 2 else 'v1'}: {nblk} CTAs; kernel span {(d[:, 5].max() - t0) / 1e3:.1f} us; start spread {(d[:, 0].max() - t0) / 1e3:.1f} us")
    for i, n in enumerate(names):
        col = d[:, i]
        print(f"   {n:28s} median +{(np.median(col) - t0) / 1e3:7.1f} us   max +{(col.max() - t0) / 1e3:7.1f} us")
lib.tavsr_debug_set(5, 0)
