"""B200 drop-in for the reference `DefaultEmbeddingLayerForAVSR`
(src/embedding_for_avsr/default.py:23-162): the per-modality embedding in front of the audio-visual
encoders.  `apply_embed_layer` and `apply_pos_enc` are separate calls because the model aligns the
two streams in time between them (src/models/avsr_espnet_model.py:427-448).

Kernel sequences:
  conv2d  (espnet Conv2dSubsamplingWOPosEnc, kernels [3,3], strides [2,2])
          conv2d_sub_im2col (conv1 + ReLU on the fly -> im2col) -> tcgen05 GEMM + ReLU (conv2)
          -> row-complete GEMM (C*F' -> d, weight columns permuted once to channels-last)
  linear  Linear + LayerNorm(eps 1e-5) as one row-complete GEMM (Dropout is the identity in eval)
  None    identity (input_size == output_size) or one row-complete GEMM
  pos-enc x * sqrt(d) by scale_add_rows; the (1, 2T-1, d) relative table is cached per length
"""
from __future__ import annotations

import math

import torch

from .. import engine, ops
from ..espnet_compat import RelPositionalEncoding
from .embedding_abs_layer import EmbeddingForAVSRAbsLayer


class Conv2dSubsamplingWOPosEnc(torch.nn.Module):
    """Parameter container with espnet's layout: conv.{0,2}, out (plain Linear)."""

    def __init__(self, idim: int, odim: int, dropout_rate: float, kernels, strides):
        super().__init__()
        if list(kernels) != [3, 3] or list(strides) != [2, 2]:
            raise NotImplementedError("the B200 conv2d embed is built for kernels [3,3], strides [2,2]")
        self.conv = torch.nn.Sequential(
            torch.nn.Conv2d(1, odim, 3, 2), torch.nn.ReLU(),
            torch.nn.Conv2d(odim, odim, 3, 2), torch.nn.ReLU())
        self.out = torch.nn.Linear(odim * (((idim - 1) // 2 - 1) // 2), odim)
        self.kernels, self.strides = list(kernels), list(strides)


class DefaultEmbeddingLayerForAVSR(EmbeddingForAVSRAbsLayer):
    """Constructor mirrors default.py:39-49."""

    def __init__(self, input_size: int, output_size: int, pos_enc_layer_type: str = "rel_pos",
                 rel_pos_type: str = "latest", input_layer: str = "conv2d", dropout_rate: float = 0.1,
                 positional_dropout_rate: float = 0.1, max_pos_emb_len: int = 5000):
        super().__init__()
        self._output_size = output_size
        self._rel_pos_type = rel_pos_type
        self._pos_enc_layer_type = pos_enc_layer_type
        if input_layer == "linear":
            self.embed = torch.nn.Sequential(torch.nn.Linear(input_size, output_size),
                                             torch.nn.LayerNorm(output_size),
                                             torch.nn.Dropout(dropout_rate))
        elif input_layer == "conv2d":
            self.embed = Conv2dSubsamplingWOPosEnc(input_size, output_size, dropout_rate,
                                                   kernels=[3, 3], strides=[2, 2])
        elif input_layer == "embed" or isinstance(input_layer, torch.nn.Module):
            raise NotImplementedError(f"input_layer={input_layer!r} is not built on the B200 path")
        elif input_layer is None:
            self.embed = None if input_size == output_size else torch.nn.Linear(input_size, output_size)
        else:
            raise ValueError("unknown input_layer: " + input_layer)
        if rel_pos_type == "legacy":
            raise NotImplementedError("rel_pos_type='legacy' is not built on the B200 path")
        elif rel_pos_type != "latest":
            raise ValueError("unknown rel_pos_type: " + rel_pos_type)
        if pos_enc_layer_type in ("abs_pos", "scaled_abs_pos", "legacy_rel_pos"):
            raise NotImplementedError(f"pos_enc_layer_type={pos_enc_layer_type!r} is not built on "
                                      "the B200 path (shipped configs use rel_pos)")
        elif pos_enc_layer_type != "rel_pos":
            raise ValueError("unknown pos_enc_layer: " + pos_enc_layer_type)
        self.pos_enc = RelPositionalEncoding(output_size, positional_dropout_rate, max_pos_emb_len)
        self._packed = engine.PackedCache()

    def output_size(self) -> int:
        return self._output_size

    # ---------------------------------------------------------------------------------------
    def apply_embed_layer(self, xs_pad: torch.Tensor, ilens: torch.Tensor):
        """(B, Tin, input_size), (B,) -> ((B, T, d), masks (B, 1, T)) (default.py:139-153)."""
        engine.require_cuda(xs_pad)
        dev = xs_pad.device
        d = self._output_size
        Tin = xs_pad.size(1)
        # ~make_pad_mask(ilens)[:, None, :] on the device: no .tolist() host sync (:141)
        masks = (torch.arange(Tin, device=dev)[None, :] < ilens.to(dev)[:, None]).unsqueeze(1)
        from .. import training
        if training.wants_grad(self, xs_pad):
            return training.avsr_embed_forward(self, xs_pad, masks)
        if self.training and any(isinstance(m, torch.nn.Dropout) and m.p > 0 for m in self.modules()):
            raise NotImplementedError("a no-grad call in train() mode with dropout enabled: the "
                                      "inference kernels have no random paths; use .eval()")
        if isinstance(self.embed, Conv2dSubsamplingWOPosEnc):
            conv, lin = self.embed.conv, self.embed.out
            B, _, Fin = xs_pad.shape
            C = conv[0].weight.shape[0]
            T, Fd = ((Tin - 1) // 2 - 1) // 2, ((Fin - 1) // 2 - 1) // 2
            pk = self._packed.get(
                "conv2d", [conv[0].weight, conv[2].weight, lin.weight],
                lambda: (conv[0].weight.reshape(C, 9).contiguous(),
                         conv[2].weight.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous(),
                         lin.weight.view(-1, C, Fd).permute(0, 2, 1).reshape(-1, Fd * C).contiguous()))
            adt = engine.act_dtype()
            a_mat = ops.conv2d_sub_im2col(xs_pad.contiguous().float(), pk[0], conv[0].bias, out_dtype=adt)
            # conv2 output fp32, last projection on TF32 operands: see encoder.py::_embed
            h2 = engine.linear(a_mat, pk[1], conv[2].bias, self._packed, "conv2", act=ops.ACT_RELU,
                               out_dtype=torch.float32).view(B * T, Fd * C)
            x = torch.empty((B * T, d), device=dev, dtype=torch.float32)
            if engine.compute_dtype() == "tf32x3":
                engine.linear_rowln(h2, pk[2], lin.bias, self._packed, "embout", out_main=x)
            else:
                ops.gemm_rowln(h2, pk[2], lin.bias, out_main=x)
            return x.view(B, T, d), masks[:, :, :-2:2][:, :, :-2:2]
        B, T, Fin = xs_pad.shape
        if self.embed is None:
            return xs_pad, masks
        x2 = engine.operand(xs_pad.reshape(B * T, Fin).contiguous().float())
        x = torch.empty((B * T, d), device=dev, dtype=torch.float32)
        if isinstance(self.embed, torch.nn.Sequential):
            lin, ln = self.embed[0], self.embed[1]
            engine.linear_rowln(x2, lin.weight, lin.bias, self._packed, "embin",
                                ln0=(ln.weight, ln.bias), eps0=ln.eps, out_main=x)
        else:
            engine.linear_rowln(x2, self.embed.weight, self.embed.bias, self._packed, "embin",
                                out_main=x)
        return x.view(B, T, d), masks

    def apply_pos_enc(self, xs_pad: torch.Tensor):
        """(B, T, d) -> ((B, T, d) * sqrt(d), pos_emb (1, 2T-1, d)) (default.py:156-162; espnet
        RelPositionalEncoding.forward, dropout is the identity in eval)."""
        engine.require_cuda(xs_pad)
        B, T, d = xs_pad.shape
        from .. import training
        if training.wants_grad(self, xs_pad):
            return training.avsr_posenc_forward(self, xs_pad)
        x2 = xs_pad.reshape(B * T, d).contiguous().float()
        sc = self._packed.get("xscale" + str(x2.device), [],
                              lambda: (torch.full((1,), math.sqrt(d), device=x2.device),
                                       torch.zeros((1,), device=x2.device)))
        y = ops.scale_add_rows(x2, x2, sc[0], sc[1], B * T)
        return y.view(B, T, d), self.pos_enc.pos_emb(T, x2.device)

    def forward(self, xs_pad: torch.Tensor, ilens: torch.Tensor):
        """(default.py:107-137): embed, then positional encoding; returns ((x, pos_emb), masks)."""
        x, masks = self.apply_embed_layer(xs_pad, ilens)
        return self.apply_pos_enc(x), masks
