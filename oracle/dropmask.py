"""TEST INFRASTRUCTURE: a deterministic dropout-mask source shared by the live reference and the
CUDA training path, so that a train-mode step with dropout active can be compared number by number.

torch's own dropout draws from the CPU generator in this container and from the CUDA generator on
the GPU box, so a seed does not carry masks across; instead both sides take their masks from
`MaskSource`: call k with shape s gets `rand(s, seed = f(seed0, k)) >= p`, scaled by 1 / (1 - p).
The reference side gets them through `patched_dropout` (torch.nn.functional.dropout replaced while
the reference's own modules run: nn.Dropout.forward and the F.dropout call at src/ctc/ctc.py:143
resolve it at call time); the CUDA side through `training.set_dropout_source`.  Equal results then
also prove that the training path visits the dropout sites in the reference's call order with the
reference's shapes."""
from __future__ import annotations

import contextlib

import torch

GOLDEN_SEED = 77   # the seed tests/golden/grad_*_dropout.npz were generated with


class MaskSource:
    def __init__(self, seed: int):
        self.seed = int(seed)
        self.calls = []          # (shape, p) in call order

    def __call__(self, shape, p: float, device=None) -> torch.Tensor:
        k = len(self.calls)
        self.calls.append((tuple(shape), float(p)))
        g = torch.Generator().manual_seed(self.seed * 100003 + k)
        keep = torch.rand(tuple(shape), generator=g) >= p
        m = keep.to(torch.float32) / (1.0 - p)
        return m if device is None else m.to(device)


@contextlib.contextmanager
def patched_dropout(source: MaskSource):
    import torch.nn.functional as F
    orig = F.dropout

    def dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:       # torch returns the input without touching the generator
            return input
        return input * source(input.shape, p).to(input.dtype)

    F.dropout = dropout
    try:
        yield source
    finally:
        F.dropout = orig
