"""Parameter containers with the espnet==202402 module/parameter names.

The reference builds its layers out of espnet leaf modules (imports at
src/encoder/branchformer/encoder.py:17-48).  The drop-in keeps exactly their attribute names,
parameter names and shapes so `state_dict()` / `load_state_dict(strict=True)` round-trip with
reference checkpoints — but these classes carry NO arithmetic: the layer
forward reads their parameters and launches the fused CUDA kernels.  When espnet2 is installed its
abstract base classes are used so `ClassChoices(type_check=AbsEncoder)` (src/tasks/asr.py:164)
accepts the drop-in.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod

import torch

try:  # pragma: no cover - espnet2 is absent in the build image
    from espnet2.asr.encoder.abs_encoder import AbsEncoder  # type: ignore
except Exception:  # noqa: BLE001
    class AbsEncoder(torch.nn.Module, ABC):
        """Stand-in for espnet2.asr.encoder.abs_encoder.AbsEncoder."""

        @abstractmethod
        def output_size(self) -> int:
            raise NotImplementedError

        @abstractmethod
        def forward(self, xs_pad, ilens, prev_states=None):
            raise NotImplementedError

try:  # pragma: no cover
    from espnet.nets.pytorch_backend.transformer.subsampling import TooShortUttError  # type: ignore
except Exception:  # noqa: BLE001
    class TooShortUttError(Exception):
        def __init__(self, message, actual_size, limit):
            super().__init__(message)
            self.actual_size = actual_size
            self.limit = limit


class _Container(torch.nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of the B200 drop-in; its arithmetic is "
            "fused into the owning encoder layer's CUDA path and cannot be called on its own")


class LayerNorm(torch.nn.LayerNorm):
    """espnet LayerNorm: torch LayerNorm with eps=1e-12."""

    def __init__(self, nout: int, dim: int = -1):
        super().__init__(nout, eps=1e-12)
        self.dim = dim


class PositionwiseFeedForward(_Container):
    def __init__(self, idim: int, hidden_units: int, dropout_rate: float, activation: str = "relu"):
        super().__init__()
        self.w_1 = torch.nn.Linear(idim, hidden_units)
        self.w_2 = torch.nn.Linear(hidden_units, idim)
        self.dropout_rate = dropout_rate
        self.activation_type = activation


class RelPositionMultiHeadedAttention(_Container):
    def __init__(self, n_head: int, n_feat: int, dropout_rate: float, zero_triu: bool = False):
        super().__init__()
        assert n_feat % n_head == 0
        self.d_k = n_feat // n_head
        self.h = n_head
        self.linear_q = torch.nn.Linear(n_feat, n_feat)
        self.linear_k = torch.nn.Linear(n_feat, n_feat)
        self.linear_v = torch.nn.Linear(n_feat, n_feat)
        self.linear_out = torch.nn.Linear(n_feat, n_feat)
        self.linear_pos = torch.nn.Linear(n_feat, n_feat, bias=False)
        self.pos_bias_u = torch.nn.Parameter(torch.empty(self.h, self.d_k))
        self.pos_bias_v = torch.nn.Parameter(torch.empty(self.h, self.d_k))
        torch.nn.init.xavier_uniform_(self.pos_bias_u)
        torch.nn.init.xavier_uniform_(self.pos_bias_v)
        self.dropout_rate = dropout_rate
        self.zero_triu = zero_triu
        self.attn = None


class ConvolutionalSpatialGatingUnit(_Container):
    def __init__(self, size: int, kernel_size: int, dropout_rate: float,
                 use_linear_after_conv: bool, gate_activation: str):
        super().__init__()
        n_channels = size // 2
        self.norm = LayerNorm(n_channels)
        self.conv = torch.nn.Conv1d(n_channels, n_channels, kernel_size, 1, (kernel_size - 1) // 2,
                                    groups=n_channels)
        self.linear = torch.nn.Linear(n_channels, n_channels) if use_linear_after_conv else None
        self.gate_activation = gate_activation
        self.kernel_size = kernel_size
        self.dropout_rate = dropout_rate

    def espnet_initialization_fn(self):
        torch.nn.init.normal_(self.conv.weight, std=1e-6)
        torch.nn.init.ones_(self.conv.bias)
        if self.linear is not None:
            torch.nn.init.normal_(self.linear.weight, std=1e-6)
            torch.nn.init.ones_(self.linear.bias)


class ConvolutionalGatingMLP(_Container):
    def __init__(self, size: int, linear_units: int, kernel_size: int, dropout_rate: float,
                 use_linear_after_conv: bool, gate_activation: str):
        super().__init__()
        self.channel_proj1 = torch.nn.Sequential(torch.nn.Linear(size, linear_units), torch.nn.GELU())
        self.csgu = ConvolutionalSpatialGatingUnit(linear_units, kernel_size, dropout_rate,
                                                   use_linear_after_conv, gate_activation)
        self.channel_proj2 = torch.nn.Linear(linear_units // 2, size)


class RelPositionalEncoding(torch.nn.Module):
    """Parameter-free; keeps a device cache of the (1, 2T-1, d) table per length."""

    def __init__(self, d_model: int, dropout_rate: float, max_len: int = 5000):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)
        self.dropout_rate = dropout_rate
        self.max_len = max_len
        self._cache = {}

    def pos_emb(self, T: int, device) -> torch.Tensor:
        key = (T, str(device))
        pe = self._cache.get(key)
        if pe is None:
            rel = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
            div = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float32)
                            * -(math.log(10000.0) / self.d_model))
            pe = torch.zeros(2 * T - 1, self.d_model)
            pe[:, 0::2] = torch.sin(rel * div)
            pe[:, 1::2] = torch.cos(rel * div)
            pe = pe.unsqueeze(0).to(device)
            if len(self._cache) > 64:
                self._cache.clear()
            self._cache[key] = pe
        return pe

    def forward(self, x: torch.Tensor):
        return x * self.xscale, self.pos_emb(x.size(1), x.device)


class Conv2dSubsampling(torch.nn.Module):
    """espnet Conv2dSubsampling parameter layout: conv.{0,2}, out.0 (Linear), out.1 (pos-enc)."""

    def __init__(self, idim: int, odim: int, dropout_rate: float, pos_enc: torch.nn.Module):
        super().__init__()
        self.conv = torch.nn.Sequential(
            torch.nn.Conv2d(1, odim, 3, 2), torch.nn.ReLU(),
            torch.nn.Conv2d(odim, odim, 3, 2), torch.nn.ReLU())
        self.out = torch.nn.Sequential(
            torch.nn.Linear(odim * (((idim - 1) // 2 - 1) // 2), odim), pos_enc)


def check_short_utt(ins, size):
    if isinstance(ins, Conv2dSubsampling) and size < 7:
        return True, 7
    return False, -1


class MultiSequential(torch.nn.Sequential):
    """espnet repeat(): Sequential whose forward threads a tuple of arguments through the layers."""

    def forward(self, *args):
        for m in self:
            args = m(*args)
        return args


def repeat(N, fn):
    return MultiSequential(*[fn(n) for n in range(N)])
