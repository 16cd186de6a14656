"""TEST INFRASTRUCTURE (oracle): espnet positional encodings restated (Appendix A.3)."""
import math

import torch


class PositionalEncoding(torch.nn.Module):
    def __init__(self, d_model, dropout_rate, max_len=5000, reverse=False):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)
        self.dropout = torch.nn.Dropout(p=dropout_rate)
        self.max_len = max_len

    def _pe(self, T, dtype, device):
        pos = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float32)
                        * -(math.log(10000.0) / self.d_model))
        pe = torch.zeros(T, self.d_model)
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        return pe.unsqueeze(0).to(device=device, dtype=dtype)

    def forward(self, x):
        return self.dropout(x * self.xscale + self._pe(x.size(1), x.dtype, x.device))


class ScaledPositionalEncoding(PositionalEncoding):
    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__(d_model, dropout_rate, max_len)
        self.alpha = torch.nn.Parameter(torch.tensor(1.0))

    def forward(self, x):
        return self.dropout(x + self.alpha * self._pe(x.size(1), x.dtype, x.device))


class RelPositionalEncoding(torch.nn.Module):
    """forward(x) -> (dropout(x*sqrt(d)), dropout(pos_emb)); pos_emb[0,k] encodes relative position
    T-1-k for k = 0..2T-2 (sin on even dims, cos on odd dims)."""

    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)
        self.dropout = torch.nn.Dropout(p=dropout_rate)
        self.max_len = max_len

    def pos_emb(self, T, dtype=torch.float32, device="cpu"):
        rel = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)  # T-1 ... -(T-1)
        div = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float32)
                        * -(math.log(10000.0) / self.d_model))
        pe = torch.zeros(2 * T - 1, self.d_model)
        pe[:, 0::2] = torch.sin(rel * div)
        pe[:, 1::2] = torch.cos(rel * div)
        return pe.unsqueeze(0).to(device=device, dtype=dtype)

    def forward(self, x):
        x = x * self.xscale
        return self.dropout(x), self.dropout(self.pos_emb(x.size(1), x.dtype, x.device))


class LegacyRelPositionalEncoding(PositionalEncoding):
    """Dormant alternative (no shipped config uses it); present so the reference imports resolve."""

    def forward(self, x):
        raise NotImplementedError("legacy_rel_pos is outside the restated surface")
