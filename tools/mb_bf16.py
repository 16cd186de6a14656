"""bf16-operand vs TF32 tiled GEMM at the C2 shapes (warm, back to back): evidence for the bf16 mode."""
import os, sys
import torch
sys.path.insert(0, os.getcwd())
from tailored_avsr_b200 import ops
DEV = "cuda"
M = 8000
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g).to(DEV)

def bench(name, fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us", flush=True)

for N, K, act, tag in ((2048, 256, 2, "channel_proj1+GELU"), (768, 256, 0, "fused QKV"),
                       (2048, 256, 1, "FFN w_1+Swish"), (256, 2048, 0, "FFN w_2 (tiled, no LN)"),
                       (256, 2304, 3, "conv2 im2col (M=37848)")):
    m = 37848 if K == 2304 else M
    x, w, b = rn(m, K), rn(N, K) / K ** 0.5, rn(N)
    out = torch.empty(m, N, device=DEV)
    bench(f"tf32 {tag} {m}x{N}x{K}", lambda: ops.gemm_bias_act(x, w, b, act=act, out=out))
    xb, wb = x.bfloat16(), w.bfloat16()
    bench(f"bf16 {tag} {m}x{N}x{K}", lambda: ops.gemm_bias_act(xb, wb, b, act=act, out=out))
