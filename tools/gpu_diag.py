"""GPU diagnostics: GEMM correctness under descriptor/tensor-map variants and kernel timings.
Run on the GPU box:  python tools/gpu_diag.py            (prints one line per experiment)
"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tailored_avsr_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n * 1e3  # us


def main():
    lib = _lib.load()
    print("device", torch.cuda.get_device_name(0), "lib version", lib.tavsr_version(), flush=True)
    g = torch.Generator().manual_seed(0)
    for tmap_tf32 in (0, 1):
        lib.tavsr_debug_set(1, tmap_tf32)
        for (M, N, K) in [(128, 256, 32), (128, 256, 256), (1000, 768, 256), (8000, 2048, 256)]:
            x = torch.randn(M, K, generator=g).to(DEV)
            w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
            b = torch.randn(N, generator=g).to(DEV)
            try:
                y = ops.gemm_bias_act(x, w, b)
                torch.cuda.synchronize()
                ref = x.double() @ w.double().t() + b.double()
                # reference with operands rounded to tf32 (rna) to separate rounding from bugs
                print(f"gemm tmap_tf32={tmap_tf32} M={M} N={N} K={K} rel_err={rel(y, ref):.3e}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"gemm tmap_tf32={tmap_tf32} M={M} N={N} K={K} FAILED: {e}", flush=True)
                return 1
    lib.tavsr_debug_set(1, 0)

    # timings vs cuBLAS TF32
    torch.backends.cuda.matmul.allow_tf32 = True
    for (M, N, K, act) in [(8000, 2048, 256, 1), (8000, 768, 256, 0), (8000, 256, 2048, 0),
                           (8000, 256, 256, 0), (8000, 256, 1024, 0), (64000, 2048, 256, 1),
                           (64000, 256, 2048, 0)]:
        x = torch.randn(M, K, device=DEV)
        w = torch.randn(N, K, device=DEV) / math.sqrt(K)
        b = torch.randn(N, device=DEV)
        out = torch.empty(M, N, device=DEV)
        if N == 256:
            main_o = torch.empty(M, 256, device=DEV)
            res = torch.randn(M, 256, device=DEV)
            gA = torch.ones(256, device=DEV)
            oA = torch.empty(M, 256, device=DEV)
            t_mine = timeit(lambda: ops.gemm_rowln(x, w, b, residual=res, alpha=0.5, out_main=main_o,
                                                   lnA=(gA, gA), out_lnA=oA))
            tag = "rowln+res+LN"
        else:
            t_mine = timeit(lambda: ops.gemm_bias_act(x, w, b, act=act, out=out))
            tag = f"tiled act={act}"
        t_ref = timeit(lambda: torch.addmm(b, x, w.t()))
        fl = 2.0 * M * N * K
        print(f"time M={M} N={N} K={K} {tag}: mine {t_mine:.1f} us ({fl / t_mine / 1e6:.1f} TF/s) "
              f"cublas-tf32 addmm {t_ref:.1f} us ({fl / t_ref / 1e6:.1f} TF/s)", flush=True)

    # other kernels at C2 shape
    B, T, H = 32, 250, 4
    M = B * T
    qkv = torch.randn(M, 768, device=DEV)
    pos = torch.randn(2 * T - 1, 256, device=DEV)
    u = torch.randn(256, device=DEV)
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
    print(f"attn B={B} T={T}: {timeit(lambda: ops.relpos_attn(qkv, pos, u, u, lens, B, T, H)):.1f} us", flush=True)
    h = torch.randn(M, 2048, device=DEV)
    ng = torch.ones(1024, device=DEV)
    cw = torch.randn(1024, 31, device=DEV)
    o = torch.empty(M, 1024, device=DEV)
    st = torch.empty(M, 2, device=DEV)
    t = timeit(lambda: ops.csgu(h, ng, ng, cw, ng, B, T, out=o, stats=st))
    print(f"csgu B={B} T={T}: {t:.1f} us  ({M * 3072 * 4 / t / 1e3:.0f} GB/s algorithmic)", flush=True)
    hs = torch.randn(M, 256, device=DEV)
    w = torch.randn(41, 256, device=DEV)
    bb = torch.randn(41, device=DEV)
    print(f"ctc_head M={M}: {timeit(lambda: ops.ctc_head(hs, w, bb, True, False, True)):.1f} us", flush=True)
    logp = torch.log_softmax(torch.randn(B, T, 41, device=DEV), -1)
    tg = torch.randint(1, 41, (B, 100), device=DEV)
    tl = torch.full((B,), 100, dtype=torch.int32, device=DEV)
    print(f"ctc_loss fwd B={B}: {timeit(lambda: ops.ctc_loss(logp, tg, lens, tl)):.1f} us", flush=True)
    print(f"ctc_loss fwd+grad B={B}: {timeit(lambda: ops.ctc_loss(logp, tg, lens, tl, want_grad=True)):.1f} us", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
