"""TEST INFRASTRUCTURE: precision-budget emulation on the CPU oracle (SURVEY.md §7 "Precision
budget").  Inside `emulate(mode)` every tensor-core contraction of the port (F.linear, matmul /
`@`) sees its two operands rounded to the tensor-core input format - "tf32": 10-bit mantissa,
round to nearest even, what TMA's TFLOAT32 maps do on load; "bf16": torch.bfloat16 - while
accumulation, LayerNorm, softmax, the residual stream and everything between kernels stay fp32:
exactly the numerics of the CUDA path's tf32 mode (and of the planned bf16-operand mode with fp32
activations).  tests/test_oracle_cpu.py pins the resulting budgets against the parity tolerances."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """Round fp32 to TF32 (1+8+10 bits), nearest-even on the dropped 13 mantissa bits."""
    if x.dtype != torch.float32:
        return x
    i = x.contiguous().view(torch.int32)
    lsb = (i >> 13) & 1
    r = (i + 0x0FFF + lsb) & ~0x1FFF
    return torch.where(torch.isfinite(x), r.view(torch.float32), x)


def _rounder(mode: str):
    if mode == "tf32":
        return round_tf32
    if mode == "bf16":
        return lambda t: t.to(torch.bfloat16).to(torch.float32) if t.dtype == torch.float32 else t
    raise ValueError(mode)


@contextlib.contextmanager
def emulate(mode: str):
    rnd = _rounder(mode)
    lin, mm, tmm = F.linear, torch.matmul, torch.Tensor.__matmul__
    F.linear = lambda x, w, b=None: lin(rnd(x), rnd(w), b)
    torch.matmul = lambda a, b: mm(rnd(a), rnd(b))
    torch.Tensor.__matmul__ = lambda a, b: mm(rnd(a), rnd(b))
    try:
        yield
    finally:
        F.linear, torch.matmul, torch.Tensor.__matmul__ = lin, mm, tmm
