"""B200 drop-in for the reference `TailoredEncoder`
(src/encoder/audiovisual/tailored/encoder.py:36-332): unified audio-visual encoder with a
per-layer, per-stream choice between rel-pos MHSA and cgMLP (heterogeneous branch widths)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .... import engine, ops
from ....espnet_compat import (ConvolutionalGatingMLP, LayerNorm, PositionwiseFeedForward,
                               RelPositionMultiHeadedAttention, repeat)
from ..audiovisual_abs_encoder import AudioVisualAbsEncoder
from .encoder_layer import TailoredEncoderLayer


class TailoredEncoder(AudioVisualAbsEncoder):
    """Constructor mirrors tailored/encoder.py:40-70."""

    def __init__(
        self,
        embed_pos_enc_layer_type,
        embed_rel_pos_type,
        output_size=256,
        attention_heads=4,
        linear_units=2048,
        num_blocks=12,
        dropout_rate=0.1,
        positional_dropout_rate=0.1,
        attention_dropout_rate=0.1,
        acoustic_branch_drop_rate=0.0,
        attention_layer_type="rel_selfattn",
        positionwise_layer_type="linear",
        ffn_activation_type="swish",
        cgmlp_linear_units=2048,
        cgmlp_conv_kernel=31,
        gate_activation="identity",
        use_linear_after_conv=False,
        acoustic_use_attn: List[bool] = [True] * 12,
        visual_use_attn: List[bool] = [False] * 12,
        macaron=True,
        zero_triu=False,
        normalize_before=True,
        ignore_id=-1,
        interctc_use_conditioning: bool = False,
        audiovisual_interctc_conditioning: bool = False,
        interctc_layer_idx: List[int] = [],
        stochastic_depth_rate=0.0,
        max_pos_emb_len: int = 5000,
    ):
        super().__init__()
        self.ignore_id = ignore_id
        self._output_size = output_size
        if embed_rel_pos_type == "legacy":
            raise NotImplementedError("embed_rel_pos_type='legacy' is not built on the B200 path")
        elif embed_rel_pos_type != "latest":
            raise ValueError("unknown embed_rel_pos_type: " + embed_rel_pos_type)
        if embed_pos_enc_layer_type in ("legacy_rel_pos", "abs_pos", "scaled_abs_pos"):
            raise NotImplementedError(f"embed_pos_enc_layer_type={embed_pos_enc_layer_type!r} is not "
                                      "built on the B200 path (shipped configs use rel_pos)")
        elif embed_pos_enc_layer_type != "rel_pos":
            raise ValueError("unknown pos_enc_layer: " + embed_pos_enc_layer_type)
        if attention_layer_type != "rel_selfattn":
            if attention_layer_type in ("selfattn", "legacy_rel_selfattn", "fast_selfattn"):
                raise NotImplementedError(f"attention_layer_type={attention_layer_type!r} is not "
                                          "built on the B200 path")
            raise ValueError("unknown attention_layer_typer: " + attention_layer_type)
        if zero_triu:
            raise NotImplementedError("zero_triu=True is not built on the B200 path")
        if positionwise_layer_type != "linear":
            raise ValueError("Support only linear.")
        engine.act_code(ffn_activation_type)

        self.normalize_before = normalize_before
        self.modality_encoding = torch.nn.Embedding(2, output_size)
        self.modality_to_id = {"audio": 0, "video": 1}

        def bcast(v, name):
            if isinstance(v, float):
                v = [v] * num_blocks
            if len(v) != num_blocks:
                raise ValueError(f"Length of {name} ({len(v)}) should be equal to num_blocks ({num_blocks})")
            return list(v)

        stochastic_depth_rate = bcast(stochastic_depth_rate, "stochastic_depth_rate")
        acoustic_branch_drop_rate = bcast(acoustic_branch_drop_rate, "acoustic_branch_drop_rate")
        assert len(acoustic_use_attn) == num_blocks, (
            f"Lenght of acoustic_use_attn ({len(acoustic_use_attn)}) should be equal to num_blocks ({num_blocks})")
        assert len(visual_use_attn) == num_blocks, (
            f"Lenght of visual_use_attn ({len(visual_use_attn)}) should be equal to num_blocks ({num_blocks})")

        def ffn():
            return PositionwiseFeedForward(output_size, linear_units, dropout_rate, ffn_activation_type)

        def attn():
            return RelPositionMultiHeadedAttention(attention_heads, output_size, attention_dropout_rate,
                                                   zero_triu)

        def cg():
            return ConvolutionalGatingMLP(output_size, cgmlp_linear_units, cgmlp_conv_kernel,
                                          dropout_rate, use_linear_after_conv, gate_activation)

        self.encoders = repeat(
            num_blocks,
            lambda lnum: TailoredEncoderLayer(
                output_size,
                ffn() if macaron else None,
                attn() if acoustic_use_attn[lnum] else None,
                cg() if not acoustic_use_attn[lnum] else None,
                attn() if visual_use_attn[lnum] else None,
                cg() if not visual_use_attn[lnum] else None,
                ffn(),
                dropout_rate,
                acoustic_branch_drop_rate[lnum],
                stochastic_depth_rate[lnum],
            ),
        )
        if self.normalize_before:
            self.after_norm = LayerNorm(output_size)
        self.interctc_layer_idx = interctc_layer_idx
        if len(interctc_layer_idx) > 0:
            assert 0 < min(interctc_layer_idx) and max(interctc_layer_idx) < num_blocks
        self.interctc_use_conditioning = interctc_use_conditioning
        self.audiovisual_interctc_conditioning = audiovisual_interctc_conditioning
        assert not (self.interctc_use_conditioning is False and self.audiovisual_interctc_conditioning is True), \
            "Audio-Visual InterCTC conditioning only can be applied if interctc_use_conditioning is set to True."
        self.conditioning_layer = None
        self._packed = engine.PackedCache()

    def output_size(self) -> int:
        return self._output_size

    def _pos_proj_all(self, pos_emb, tag: str):
        layers = [(i, getattr(l, tag + "_attn")) for i, l in enumerate(self.encoders)
                  if getattr(l, tag + "_attn") is not None]
        if not layers or pos_emb is None:
            return {}
        d = self._output_size
        ws = [a.linear_pos.weight for _, a in layers]
        wcat = self._packed.get("wpos_" + tag, ws, lambda: torch.cat(ws, 0).contiguous())
        pe = engine.operand(pos_emb.reshape(-1, d).contiguous().float())
        p_all = engine.linear(pe, wcat, None, self._packed, "wpos_" + tag + ".w",
                              out_dtype=engine.act_dtype())
        return {i: p_all[:, k * d:(k + 1) * d] for k, (i, _) in enumerate(layers)}

    def forward(self, audio_pad, audio_masks, video_pad, video_masks, prev_states=None, ctc=None,
                audiovisual_fusion=None):
        """Same contract as the reference forward (tailored/encoder.py:221-249)."""
        taps_on = len(self.interctc_layer_idx) > 0
        if taps_on:
            if audiovisual_fusion is None or not hasattr(audiovisual_fusion, "run"):
                raise ValueError("audio-visual InterCTC taps need the B200 `audiovisual_fusion` "
                                 "module (tailored/encoder.py:280-286)")
            if self.interctc_use_conditioning and (ctc is None or self.conditioning_layer is None):
                raise ValueError("InterCTC self-conditioning needs the `ctc` module and an assigned "
                                 "`conditioning_layer` (avsr_espnet_model.py)")
        audio, a_pos = audio_pad if isinstance(audio_pad, tuple) else (audio_pad, None)
        video, v_pos = video_pad if isinstance(video_pad, tuple) else (video_pad, None)
        engine.require_cuda(audio, video)
        from .... import training
        if training.wants_grad(self, audio, video):
            return training.tailored_encoder_forward(self, audio_pad, audio_masks, video_pad, video_masks,
                                                     ctc=ctc, fusion=audiovisual_fusion)
        if audio.shape != video.shape:
            raise NotImplementedError("the B200 tailored encoder expects time-aligned streams of "
                                      "equal shape (avsr_espnet_model.py:439 aligns them)")
        B, T, d = audio.shape
        M = B * T
        me = self.modality_encoding.weight
        # modality encoding (:251-263) while stacking the two streams into one (2*B*T, d) matrix
        x = torch.cat([(audio + me[0]).reshape(M, d), (video + me[1]).reshape(M, d)], 0).contiguous().float()
        first = self.encoders[0]
        xn = ops.layernorm(x, first.norm_ff_macaron.weight, first.norm_ff_macaron.bias, eps=1e-12,
                           out_dtype=engine.act_dtype())
        la = engine.lens_from_mask(audio_masks, B, T, x.device)
        lv = engine.lens_from_mask(video_masks, B, T, x.device)
        pa = self._pos_proj_all(a_pos, "acoustic")
        pv = self._pos_proj_all(v_pos, "visual")
        n = len(self.encoders)
        inter = []
        after = (self.after_norm.weight, self.after_norm.bias) if self.normalize_before else None
        for i, layer in enumerate(self.encoders):
            layer._check_supported()
            if i + 1 < n:
                nxt = self.encoders[i + 1]
                next_norm = (nxt.norm_ff_macaron.weight, nxt.norm_ff_macaron.bias)
            else:
                next_norm = after
            # the last block's trailing LayerNorm is after_norm: the fp32 encoder output
            nn_dtype = torch.float32 if i + 1 == n else None
            x, xn = layer.run(x, xn, pa.get(i), pv.get(i), la, lv, B, T, next_norm=next_norm,
                              next_norm_dtype=nn_dtype)
            if taps_on and (i + 1) in self.interctc_layer_idx:
                # intermediate outputs are normalised too (:275-278), fused (:280-286), and with
                # conditioning the posteriors are fed back into BOTH streams (:291-318)
                if self.normalize_before:
                    t_out = xn if i + 1 == n else ops.layernorm(x, after[0], after[1], eps=1e-12)
                else:
                    t_out = x
                fused = audiovisual_fusion.run(t_out[:M], t_out[M:], la, lv, B, T)
                inter.append((i + 1, fused.view(B, T, d)))
                if self.interctc_use_conditioning:
                    cl = self.conditioning_layer
                    if self.audiovisual_interctc_conditioning:
                        p_av = ctc.softmax(fused.view(B, T, d)).reshape(M, -1)
                        prob = torch.cat([p_av, p_av], 0)
                    else:
                        prob = ctc.softmax(t_out.view(2 * B, T, d)).reshape(2 * M, -1)
                    x, xn = ops.vocab_residual(x, prob.contiguous(), cl.weight.contiguous(), cl.bias,
                                               ln=next_norm, ln_dtype=nn_dtype or engine.act_dtype())
        out = xn if self.normalize_before else x
        a_out = out[:M].view(B, T, d)
        if inter:
            return (a_out, inter), audio_masks, out[M:].view(B, T, d), video_masks, None
        return a_out, audio_masks, out[M:].view(B, T, d), video_masks, None
