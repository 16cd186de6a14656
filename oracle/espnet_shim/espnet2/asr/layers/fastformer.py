"""TEST INFRASTRUCTURE (oracle): import stub (only used in isinstance checks by the reference)."""
import torch


class FastSelfAttention(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("fast_selfattn is outside the restated surface")
