"""TEST INFRASTRUCTURE: deterministic synthetic weights and inputs shared by the oracle, the golden
generator and the CUDA parity tests.  Weights are a pure function of the parameter
NAME and shape, so the reference modules (here), the oracle and the CUDA drop-in (on the GPU box,
where /root/reference does not exist) can all be filled identically without shipping tensors."""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    leaf = name.rsplit(".", 1)[-1]
    if len(shape) == 1:
        is_norm_scale = leaf == "weight"
        return 1.0 + 0.1 * t if is_norm_scale else 0.1 * t
    if leaf in ("pos_bias_u", "pos_bias_v"):
        return 0.2 * t
    if "modality_encoding" in name:
        return 0.5 * t
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return t / math.sqrt(fan_in)


def synth_state_dict(named_shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0
                     ) -> Dict[str, torch.Tensor]:
    return {n: synth_tensor(n, tuple(s), seed) for n, s in named_shapes}


def fill_module(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    """Overwrite every parameter of `module` with synth_tensor(prefix + name); returns the dict."""
    sd = {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            t = synth_tensor(prefix + name, tuple(p.shape), seed)
            p.copy_(t)
            sd[prefix + name] = t
    return sd


def randn(shape, seed: int) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def rand_lens(B: int, T: int, seed: int, lo: float = 0.4) -> torch.Tensor:
    """lens = floor(T * U[lo,1]) with at least one full-length item."""
    g = torch.Generator().manual_seed(seed)
    lens = (T * (lo + (1 - lo) * torch.rand(B, generator=g))).floor().long().clamp(1, T)
    lens[0] = T
    return lens


def rand_targets(B: int, Lmax: int, V: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(1, V - 1, (B, Lmax), generator=g)
