"""TEST INFRASTRUCTURE: deterministic synthetic weights and inputs shared by the oracle, the golden
generator and the CUDA parity tests.  Weights are a pure function of the parameter
NAME and shape, so the reference modules (here), the oracle and the CUDA drop-in (on the GPU box,
where /root/reference does not exist) can all be filled identically without shipping tensors."""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0, hot: bool = False) -> torch.Tensor:
    """Default-init-like values ("weights = module default init"): matrices, conv
    taps and their biases ~ U(-1/sqrt(fan_in), +1/sqrt(fan_in)) like torch's Linear/Conv defaults,
    pos_bias_u/v ~ xavier-uniform range, embeddings ~ N(0,1); LayerNorm affines are perturbed
    (1 + 0.1 n, 0.1 n) so the affine paths are exercised.  hot=True draws N(0, 1/fan_in) matrices
    (1.7x larger branch outputs): the TF32 error-growth stress case."""
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    is_norm = "norm" in parent or parent in ("1",) and len(shape) == 1  # embed.1 = LayerNorm
    if len(shape) == 1 and (is_norm or hot):
        t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
        return 1.0 + 0.1 * t if leaf == "weight" else 0.1 * t
    if leaf in ("pos_bias_u", "pos_bias_v"):
        t = torch.rand(tuple(shape), generator=g, dtype=torch.float32) * 2 - 1
        return t * math.sqrt(6.0 / (shape[0] + shape[1]))
    if "modality_encoding" in name:
        return torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if len(shape) == 1:
        # bias of a Linear/Conv: bound 1/sqrt(fan_in); fan_in is not known from the bias alone,
        # the width of the model (256) is a representative stand-in
        t = torch.rand(tuple(shape), generator=g, dtype=torch.float32) * 2 - 1
        return t / 16.0
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    if hot:
        return torch.randn(tuple(shape), generator=g, dtype=torch.float32) / math.sqrt(fan_in)
    t = torch.rand(tuple(shape), generator=g, dtype=torch.float32) * 2 - 1
    return t / math.sqrt(fan_in)


def synth_state_dict(named_shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0,
                     hot: bool = False) -> Dict[str, torch.Tensor]:
    return {n: synth_tensor(n, tuple(s), seed, hot) for n, s in named_shapes}


def fill_module(module: torch.nn.Module, seed: int = 0, prefix: str = "", hot: bool = False
                ) -> Dict[str, torch.Tensor]:
    """Overwrite every parameter of `module` with synth_tensor(prefix + name); returns the dict."""
    sd = {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            t = synth_tensor(prefix + name, tuple(p.shape), seed, hot)
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
