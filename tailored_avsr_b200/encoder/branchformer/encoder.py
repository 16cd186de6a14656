"""B200 drop-in for the reference `MyBranchformerEncoder`
(src/encoder/branchformer/encoder.py:52-412): same constructor kwargs, attribute surface
(`encoders`, `embed`, `after_norm`, `normalize_before`, `interctc_*`, `conditioning_layer`),
`state_dict` layout and forward contract; the layer stack runs as fused sm_100a kernels."""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from ... import engine, ops, training
from ...espnet_compat import (AbsEncoder, Conv2dSubsampling, ConvolutionalGatingMLP, LayerNorm,
                              PositionwiseFeedForward, RelPositionalEncoding,
                              RelPositionMultiHeadedAttention, TooShortUttError, check_short_utt,
                              repeat)
from .encoder_layer import MyBranchformerEncoderLayer

_UNBUILT_INPUT_LAYERS = ("conv1d", "conv3dresnet18", "conv1d2", "conv1d3", "conv2d1", "conv2d2",
                         "conv2d6", "conv2d8", "embed")


def _broadcast(value, n: int, name: str) -> List[float]:
    if isinstance(value, float):
        value = [value] * n
    if len(value) != n:
        raise ValueError(f"Length of {name} ({len(value)}) should be equal to num_blocks ({n})")
    return list(value)


class MyBranchformerEncoder(AbsEncoder):
    """Branchformer encoder (B200 path).  Constructor mirrors encoder.py:56-89."""

    def __init__(
        self,
        input_size=256,
        output_size=256,
        attention_heads=4,
        linear_units=2048,
        num_blocks=6,
        cgmlp_linear_units=2048,
        cgmlp_conv_kernel=31,
        cgmlp_weight=0.5,
        dropout_rate=0.1,
        positional_dropout_rate=0.1,
        attention_dropout_rate=0.1,
        attn_branch_drop_rate=0.0,
        input_layer="conv3dresnet18",
        rel_pos_type="latest",
        pos_enc_layer_type="rel_pos",
        attention_layer_type="rel_selfattn",
        positionwise_layer_type="linear",
        ffn_activation_type="relu",
        merge_method="learned_ave",
        gate_activation="identity",
        ignore_id=-1,
        use_attn=True,
        use_cgmlp=True,
        macaron=True,
        zero_triu=False,
        normalize_before=True,
        use_linear_after_conv=False,
        interctc_use_conditioning: bool = False,
        interctc_layer_idx: List[int] = [],
        stochastic_depth_rate=0.0,
        max_pos_emb_len: int = 5000,
    ):
        super().__init__()
        self._output_size = output_size

        if rel_pos_type == "legacy":
            raise NotImplementedError("rel_pos_type='legacy' is not built on the B200 path")
        elif rel_pos_type != "latest":
            raise ValueError("unknown rel_pos_type: " + rel_pos_type)
        if pos_enc_layer_type in ("abs_pos", "scaled_abs_pos", "legacy_rel_pos"):
            raise NotImplementedError(f"pos_enc_layer_type={pos_enc_layer_type!r} is not built on "
                                      "the B200 path (shipped configs use rel_pos)")
        elif pos_enc_layer_type != "rel_pos":
            raise ValueError("unknown pos_enc_layer: " + pos_enc_layer_type)
        if attention_layer_type in ("selfattn", "legacy_rel_selfattn", "fast_selfattn"):
            raise NotImplementedError(f"attention_layer_type={attention_layer_type!r} is not built "
                                      "on the B200 path (shipped configs use rel_selfattn)")
        elif attention_layer_type != "rel_selfattn":
            raise ValueError("unknown encoder_attn_layer: " + attention_layer_type)
        if zero_triu:
            raise NotImplementedError("zero_triu=True is not built on the B200 path")

        def pos_enc():
            return RelPositionalEncoding(output_size, positional_dropout_rate, max_pos_emb_len)

        # -- embedding layer (encoder.py:122-203)
        if input_layer == "linear":
            self.embed = torch.nn.Sequential(
                torch.nn.Linear(input_size, output_size),
                torch.nn.LayerNorm(output_size),
                torch.nn.Dropout(dropout_rate),
                pos_enc(),
            )
        elif input_layer == "conv2d":
            self.embed = Conv2dSubsampling(input_size, output_size, dropout_rate, pos_enc())
        elif input_layer is None:
            self.embed = None
        elif isinstance(input_layer, str) and input_layer in _UNBUILT_INPUT_LAYERS:
            raise NotImplementedError(f"input_layer={input_layer!r} is not built on the B200 path "
                                      "(shipped configs use conv2d, linear or None)")
        else:
            raise ValueError("unknown input_layer: " + str(input_layer))
        self.input_layer = input_layer

        self.normalize_before = normalize_before
        if positionwise_layer_type != "linear":
            raise ValueError("Support only linear.")
        engine.act_code(ffn_activation_type)  # validates

        stochastic_depth_rate = _broadcast(stochastic_depth_rate, num_blocks, "stochastic_depth_rate")
        cgmlp_weight = _broadcast(cgmlp_weight, num_blocks, "cgmlp_weight")
        attn_branch_drop_rate = _broadcast(attn_branch_drop_rate, num_blocks, "attn_branch_drop_rate")

        def ffn():
            return PositionwiseFeedForward(output_size, linear_units, dropout_rate, ffn_activation_type)

        self.encoders = repeat(
            num_blocks,
            lambda lnum: MyBranchformerEncoderLayer(
                output_size,
                RelPositionMultiHeadedAttention(attention_heads, output_size, attention_dropout_rate,
                                                zero_triu) if use_attn else None,
                ConvolutionalGatingMLP(output_size, cgmlp_linear_units, cgmlp_conv_kernel,
                                       dropout_rate, use_linear_after_conv, gate_activation)
                if use_cgmlp else None,
                ffn() if macaron else None,
                ffn(),
                dropout_rate,
                merge_method,
                cgmlp_weight[lnum],
                attn_branch_drop_rate[lnum],
                stochastic_depth_rate[lnum],
            ),
        )
        if self.normalize_before:
            self.after_norm = LayerNorm(output_size)

        self.interctc_layer_idx = interctc_layer_idx
        if len(interctc_layer_idx) > 0:
            assert 0 < min(interctc_layer_idx) and max(interctc_layer_idx) < num_blocks
        self.interctc_use_conditioning = interctc_use_conditioning
        self.conditioning_layer = None
        self._packed = engine.PackedCache()

    def output_size(self) -> int:
        return self._output_size

    # ---------------------------------------------------------------------------------------
    def _pos_proj_all(self, pos_emb: torch.Tensor):
        """linear_pos of every block in ONE GEMM: (2T-1, n_attn*d); returns per-layer views."""
        attn_layers = [(i, l.attn) for i, l in enumerate(self.encoders) if l.attn is not None]
        if not attn_layers:
            return {}
        d = self._output_size
        ws = [a.linear_pos.weight for _, a in attn_layers]
        wcat = self._packed.get("wpos", ws, lambda: torch.cat(ws, 0).contiguous())
        pe = engine.operand(pos_emb.reshape(-1, d).contiguous().float())
        p_all = engine.linear(pe, wcat, None, self._packed, "wpos.w", out_dtype=engine.act_dtype())
        return {i: p_all[:, k * d:(k + 1) * d] for k, (i, _) in enumerate(attn_layers)}

    def _embed(self, xs_pad: torch.Tensor, masks: torch.Tensor, first_norm):
        """Input layer + x*sqrt(d) + first LayerNorm.  Returns (x2d, xn2d, pos_emb, masks, B, T);
        x2d is the fp32 residual stream, xn2d the first block's LayerNorm in operand storage."""
        d = self._output_size
        adt = engine.act_dtype()
        if isinstance(self.embed, Conv2dSubsampling):
            short_status, limit_size = check_short_utt(self.embed, xs_pad.size(1))
            if short_status:
                raise TooShortUttError(
                    f"has {xs_pad.size(1)} frames and is too short for subsampling "
                    + f"(it needs more than {limit_size} frames), return empty results",
                    xs_pad.size(1), limit_size)
            # conv1 + ReLU are evaluated inside the im2col writer, conv2 + ReLU is one tcgen05 GEMM
            # over K = 9 C, and the 4864 -> 256 projection, the sqrt(d) scale and the first
            # LayerNorm are one row-complete GEMM on the channels-last result (weights permuted
            # once): no cuDNN / ATen arithmetic on the path.  TAVSR_CUDNN_EMBED=1 keeps the former
            # cuDNN convolutions as a cross-check.
            conv = self.embed.conv
            lin = self.embed.out[0]
            B, Tin, Fin = xs_pad.shape
            C = conv[0].weight.shape[0]
            T, Fd = ((Tin - 1) // 2 - 1) // 2, ((Fin - 1) // 2 - 1) // 2
            x = torch.empty((B * T, d), device=xs_pad.device, dtype=torch.float32)
            xn = torch.empty((B * T, d), device=xs_pad.device, dtype=adt)
            if engine.CUDNN_EMBED:
                h = F.relu(F.conv2d(xs_pad.unsqueeze(1), conv[0].weight, conv[0].bias, stride=2))
                h = F.relu(F.conv2d(h, conv[2].weight, conv[2].bias, stride=2))
                h2 = engine.operand(h.transpose(1, 2).contiguous().view(B * T, C * Fd))
                w_lin = lin.weight
            else:
                pk = self._packed.get(
                    "conv2d", [conv[0].weight, conv[2].weight, lin.weight],
                    lambda: (conv[0].weight.reshape(C, 9).contiguous(),
                             conv[2].weight.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous(),
                             lin.weight.view(-1, C, Fd).permute(0, 2, 1).reshape(-1, Fd * C).contiguous()))
                # bf16 mode: the im2col operand (the big tensor: 9 C values per conv2 input) and the
                # conv2 weights are bf16, but conv2's output stays fp32 and the 4864 -> 256
                # projection that PRODUCES the residual stream runs on TF32 operands - an error
                # made here rides the residual path through every block unattenuated
                a_mat = ops.conv2d_sub_im2col(xs_pad.contiguous().float(), pk[0], conv[0].bias,
                                              out_dtype=adt)
                h2 = engine.linear(a_mat, pk[1], conv[2].bias, self._packed, "conv2", act=ops.ACT_RELU,
                                   out_dtype=torch.float32).view(B * T, Fd * C)
                w_lin = pk[2]
            if engine.compute_dtype() == "tf32x3":
                engine.linear_rowln(h2, w_lin, lin.bias, self._packed, "embout", alpha=math.sqrt(d),
                                    out_main=x, lnA=first_norm, out_lnA=xn)
            elif adt == torch.float32:
                ops.gemm_rowln(h2, w_lin, lin.bias, alpha=math.sqrt(d), out_main=x, lnA=first_norm,
                               out_lnA=xn)
            else:   # the TF32 row-complete kernel stores fp32 only: LayerNorm -> bf16 as its own pass
                ops.gemm_rowln(h2.float() if h2.dtype != torch.float32 else h2, w_lin, lin.bias,
                               alpha=math.sqrt(d), out_main=x)
                ops.layernorm(x, first_norm[0], first_norm[1], eps=1e-12, out=xn)
            masks = masks[:, :, :-2:2][:, :, :-2:2]
            pos_emb = self.embed.out[1].pos_emb(T, x.device)
        elif self.embed is not None:
            B, T, Fd = xs_pad.shape
            lin, ln = self.embed[0], self.embed[1]
            sc = math.sqrt(d)
            g16, b16 = self._packed.get("embln", [ln.weight, ln.bias],
                                        lambda: ((ln.weight * sc).contiguous(), (ln.bias * sc).contiguous()))
            x = torch.empty((B * T, d), device=xs_pad.device, dtype=torch.float32)
            xn = torch.empty((B * T, d), device=xs_pad.device, dtype=adt)
            xin = engine.operand(xs_pad.reshape(B * T, Fd).contiguous().float())
            engine.linear_rowln(xin, lin.weight, lin.bias, self._packed, "embin",
                                ln0=(g16, b16), eps0=ln.eps, out_main=x, lnA=first_norm, out_lnA=xn)
            pos_emb = self.embed[3].pos_emb(T, x.device)
        else:
            if isinstance(xs_pad, tuple):
                xs, pos_emb = xs_pad
            else:
                raise NotImplementedError("input_layer=None expects (x, pos_emb) like the AV wrappers "
                                          "pass (conventional/encoder.py:149)")
            B, T, _ = xs.shape
            x = xs.reshape(B * T, d).contiguous().float()
            xn = ops.layernorm(x, first_norm[0], first_norm[1], eps=1e-12, out_dtype=adt)
        return x, xn, pos_emb, masks, B, T

    def run_blocks(self, x, xn, pos_emb, lens, B, T, taps=(), stop_after: Optional[int] = None,
                   ctc=None):
        """The block stack on 2-D activations.  Returns (out, tap_outputs) where `out` already went
        through after_norm when normalize_before.  With `interctc_use_conditioning` the tapped
        posteriors are fed back, x += conditioning_layer(ctc.softmax(tap)) (encoder.py:393-401)."""
        pos = self._pos_proj_all(pos_emb)
        n = len(self.encoders)
        last = n - 1 if stop_after is None else min(stop_after, n - 1)
        after = (self.after_norm.weight, self.after_norm.bias) if self.normalize_before else None
        tap_outs = []
        for i, layer in enumerate(self.encoders):
            if i > last:
                break
            layer._check_supported()
            if i < last:
                nxt = self.encoders[i + 1]
                next_norm = (nxt.norm_ff_macaron.weight, nxt.norm_ff_macaron.bias)
            else:
                next_norm = after
            # the last block's trailing LayerNorm is after_norm, i.e. the fp32 encoder output
            nn_dtype = torch.float32 if i == last else None
            y, yn = layer.run(x, xn, pos.get(i), lens, B, T, next_norm=next_norm,
                              next_norm_dtype=nn_dtype)
            if (i + 1) in taps:
                t_out = y
                if self.normalize_before:
                    t_out = yn if i == last else ops.layernorm(y, after[0], after[1], eps=1e-12)
                tap_outs.append((i + 1, t_out.view(B, T, -1)))
                if self.interctc_use_conditioning:
                    prob = ctc.softmax(t_out.view(B, T, -1))
                    cl = self.conditioning_layer
                    y, yn = ops.vocab_residual(y, prob.reshape(B * T, -1).contiguous(),
                                               cl.weight.contiguous(), cl.bias, ln=next_norm,
                                               ln_dtype=nn_dtype or engine.act_dtype())
            x, xn = y, yn
        out = xn if self.normalize_before else x
        return out, tap_outs

    def forward(self, xs_pad: torch.Tensor, ilens: torch.Tensor, prev_states: torch.Tensor = None,
                ctc=None, max_layer: int = None
                ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """Same contract as the reference forward (encoder.py:324-343)."""
        x_in = xs_pad[0] if isinstance(xs_pad, tuple) else xs_pad
        engine.require_cuda(x_in)
        if training.wants_grad(self, x_in):
            # training step: one autograd node per block on the backward kernels (training.py)
            return training.encoder_forward(self, xs_pad, ilens, max_layer=max_layer, ctc=ctc)
        if self.interctc_use_conditioning and len(self.interctc_layer_idx) > 0:
            if ctc is None or self.conditioning_layer is None:
                raise ValueError("InterCTC self-conditioning needs the `ctc` module and an assigned "
                                 "`conditioning_layer` (espnet_model.py:106-112)")
        Tin = x_in.size(1)
        dev = x_in.device
        # ~make_pad_mask(ilens)[:, None, :] built on the device: no .tolist() host sync (:345)
        masks = (torch.arange(Tin, device=dev)[None, :] < ilens.to(dev)[:, None]).unsqueeze(1)
        first = self.encoders[0]
        first_norm = (first.norm_ff_macaron.weight, first.norm_ff_macaron.bias)
        x, xn, pos_emb, masks, B, T = self._embed(xs_pad, masks, first_norm)
        lens = masks.reshape(B, -1).sum(dim=1).to(torch.int32)
        stop = None
        if len(self.interctc_layer_idx) == 0 and max_layer is not None and 0 <= max_layer < len(self.encoders):
            stop = max_layer
        out, taps = self.run_blocks(x, xn, pos_emb, lens, B, T, taps=tuple(self.interctc_layer_idx),
                                    stop_after=stop, ctc=ctc)
        out = out.view(B, T, self._output_size)
        olens = masks.squeeze(1).sum(1)
        if len(taps) > 0:
            return (out, taps), olens, None
        return out, olens, None
