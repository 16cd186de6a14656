"""Debug aid for the tcgen05 attention kernel: compares probabilities (debug mode 1) and the
context against an fp64 reference on one small case."""
import math
import sys

import torch

sys.path.insert(0, ".")
from tailored_avsr_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


def rel_shift(x):
    b, h, t, n = x.shape
    zp = torch.zeros((b, h, t, 1), dtype=x.dtype)
    xp = torch.cat([zp, x], dim=-1).view(b, h, n + 1, t)
    return xp[:, :, 1:].view_as(x)[:, :, :, : n // 2 + 1]


def main(B=1, T=64, lens=(64,)):
    H, dk = 4, 64
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B * T, 3 * H * dk, generator=g)
    pos = torch.randn(2 * T - 1, H * dk, generator=g)
    u = torch.randn(H * dk, generator=g) * 0.5
    v = torch.randn(H * dk, generator=g) * 0.5
    lens_t = torch.tensor(lens, dtype=torch.int32)
    q, k, vv = [t.double().view(B, T, H, dk).transpose(1, 2) for t in qkv.split(H * dk, dim=1)]
    p = pos.double().view(1, 2 * T - 1, H, dk).transpose(1, 2)
    ac = (q + u.double().view(1, H, 1, dk)) @ k.transpose(-2, -1)
    bd = rel_shift((q + v.double().view(1, H, 1, dk)) @ p.transpose(-2, -1))
    scores = (ac + bd) / math.sqrt(dk)
    mask = (torch.arange(T)[None, :] >= lens_t[:, None].long())[:, None, None, :]
    scores = scores.masked_fill(mask, torch.finfo(torch.float64).min)
    attn = torch.softmax(scores, dim=-1).masked_fill(mask, 0.0)
    ref = (attn @ vv).transpose(1, 2).reshape(B * T, H * dk)
    args = (qkv.to(DEV), pos.to(DEV), u.to(DEV), v.to(DEV), lens_t.to(DEV), B, T, H)
    for mode in (1, 0):
        _lib.load().tavsr_debug_set(9, mode)
        out = ops.relpos_attn(*args, round_out=False).cpu().double()
        torch.cuda.synchronize()
        if mode == 1:
            want = attn[:, :, :, :64].transpose(1, 2).reshape(B * T, H * 64)
            if T < 64:
                want = torch.nn.functional.pad(attn, (0, 64 - T)).transpose(1, 2).reshape(B * T, H * 64)
            err = (out - want).abs().max().item()
            print(f"probabilities: max|d| {err:.3e}  got[0,:4] {out[0,:4].tolist()} want {want[0,:4].tolist()}")
        else:
            err = (out - ref).norm() / ref.norm()
            print(f"context: rel fro {err:.3e}  got[0,:4] {out[0,:4].tolist()} want {ref[0,:4].tolist()}")
    _lib.load().tavsr_debug_set(9, 0)


if __name__ == "__main__":
    main()
    main(2, 250, (250, 130))
