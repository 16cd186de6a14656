"""Time the rel-pos attention kernel at the C2 shape in both compute modes and print the in-kernel
phase stamps of the tf32 build (usage: python tools/time_attn.py)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from tailored_avsr_b200 import _lib, ops  # noqa: E402

B, T, H = 32, 250, 4
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).cuda()
qkv, pos = rn(B * T, 768), rn(2 * T - 1, 3072)
u, v = rn(256) * 0.5, rn(256) * 0.5
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
lib = _lib.load()


qkv16, pos16 = qkv.bfloat16(), pos.bfloat16()


def attn():
    return ops.relpos_attn(qkv, pos[:, 256:512], u, v, lens, B, T, H, round_out=True)


def attn16():
    return ops.relpos_attn(qkv16, pos16[:, 256:512], u, v, lens, B, T, H)


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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


print("tf32 attention us", t(attn))
print("bf16 attention us", t(attn16))
nblk = 2 * H * B
dbg = torch.zeros(nblk * 16, dtype=torch.int64, device="cuda")
lib.tavsr_debug_set_ptr(dbg.data_ptr())
attn(); torch.cuda.synchronize()
dbg.zero_()
attn(); torch.cuda.synchronize()
lib.tavsr_debug_set_ptr(None)
d = dbg.cpu().numpy().reshape(nblk, 16).astype(np.float64)
t0 = d[:, 0].min()
names = {0: "start", 1: "t0 k_full", 2: "t0 cu/cv done", 3: "t0 s_done", 4: "t0 pass1", 5: "t0 pass2+rescale",
         6: "t0 transpose/p_ready", 7: "t1 k_full", 8: "t1 cu/cv done", 9: "t1 s_done", 10: "t1 pass1",
         11: "t1 pass2+rescale", 12: "t1 transpose/p_ready", 13: "o_done", 14: "end"}
print(f"{nblk} CTAs; kernel span {(d[:, 14].max() - t0) / 1e3:.1f} us; start spread {(d[:, 0].max() - t0) / 1e3:.1f} us")
rel = d - d[:, :1]
for i, n in names.items():
    print(f"  {n:24s} median +{np.median(rel[:, i]) / 1e3:6.2f} us   max +{rel[:, i].max() / 1e3:6.2f} us")
