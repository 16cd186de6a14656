"""TEST INFRASTRUCTURE (oracle): espnet2 ConvolutionalGatingMLP restated (Appendix A.6)."""
import torch

from espnet.nets.pytorch_backend.nets_utils import get_activation
from espnet.nets.pytorch_backend.transformer.layer_norm import LayerNorm


class ConvolutionalSpatialGatingUnit(torch.nn.Module):
    def __init__(self, size, kernel_size, dropout_rate, use_linear_after_conv, gate_activation):
        super().__init__()
        n_channels = size // 2
        self.norm = LayerNorm(n_channels)
        self.conv = torch.nn.Conv1d(n_channels, n_channels, kernel_size, 1,
                                    (kernel_size - 1) // 2, groups=n_channels)
        self.linear = torch.nn.Linear(n_channels, n_channels) if use_linear_after_conv else None
        self.act = torch.nn.Identity() if gate_activation == "identity" \
            else get_activation(gate_activation)
        self.dropout = torch.nn.Dropout(dropout_rate)

    def espnet_initialization_fn(self):
        torch.nn.init.normal_(self.conv.weight, std=1e-6)
        torch.nn.init.ones_(self.conv.bias)
        if self.linear is not None:
            torch.nn.init.normal_(self.linear.weight, std=1e-6)
            torch.nn.init.ones_(self.linear.bias)

    def forward(self, x, gate_add=None):
        x_r, x_g = x.chunk(2, dim=-1)
        x_g = self.norm(x_g)
        x_g = self.conv(x_g.transpose(1, 2)).transpose(1, 2)
        if self.linear is not None:
            x_g = self.linear(x_g)
        if gate_add is not None:
            x_g = x_g + gate_add
        x_g = self.act(x_g)
        return self.dropout(x_r * x_g)


class ConvolutionalGatingMLP(torch.nn.Module):
    def __init__(self, size, linear_units, kernel_size, dropout_rate, use_linear_after_conv,
                 gate_activation):
        super().__init__()
        self.channel_proj1 = torch.nn.Sequential(torch.nn.Linear(size, linear_units),
                                                 torch.nn.GELU())
        self.csgu = ConvolutionalSpatialGatingUnit(linear_units, kernel_size, dropout_rate,
                                                   use_linear_after_conv, gate_activation)
        self.channel_proj2 = torch.nn.Linear(linear_units // 2, size)

    def forward(self, x, mask):
        if isinstance(x, tuple):
            xs_pad, pos_emb = x
        else:
            xs_pad, pos_emb = x, None
        xs_pad = self.channel_proj1(xs_pad)
        xs_pad = self.csgu(xs_pad)
        xs_pad = self.channel_proj2(xs_pad)
        return (xs_pad, pos_emb) if pos_emb is not None else xs_pad
