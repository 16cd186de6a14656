"""TEST INFRASTRUCTURE (oracle): espnet 202402 Conv2dSubsamplingWOPosEnc restated from its published
behaviour (imported by the reference at src/embedding_for_avsr/default.py:19): a stack of
Conv2d(k, stride s) + ReLU over (B, 1, T, idim), then Linear(odim * olen, odim); no positional
encoding; the mask is sub-sampled with x_mask[:, :, : -k + 1 : s] per layer."""
import math

import torch


class Conv2dSubsamplingWOPosEnc(torch.nn.Module):
    def __init__(self, idim, odim, dropout_rate, kernels, strides):
        assert len(kernels) == len(strides)
        super().__init__()
        conv = []
        olen = idim
        for i, (k, s) in enumerate(zip(kernels, strides)):
            conv += [torch.nn.Conv2d(1 if i == 0 else odim, odim, k, s), torch.nn.ReLU()]
            olen = math.floor((olen - k) / s + 1)
        self.conv = torch.nn.Sequential(*conv)
        self.out = torch.nn.Linear(odim * olen, odim)
        self.strides = strides
        self.kernels = kernels

    def forward(self, x, x_mask):
        x = x.unsqueeze(1)  # (b, c, t, f)
        x = self.conv(x)
        b, c, t, f = x.size()
        x = self.out(x.transpose(1, 2).contiguous().view(b, t, c * f))
        if x_mask is None:
            return x, None
        for k, s in zip(self.kernels, self.strides):
            x_mask = x_mask[:, :, : -k + 1 : s]
        return x, x_mask
