"""B200 drop-in for the reference `AdaptiveAudioVisualFusion`
(src/audiovisual_fusion/adaptive_audiovisual_fusion.py:29-211): same constructor, parameter names,
`acoustic_weight` / `visual_weight` attributes and forward contract.  It sits between the
audio-visual encoder and CTC (src/models/avsr_espnet_model.py:467) and inside the tailored
encoder's InterCTC taps (tailored/encoder.py:280-289).

Kernel sequence (learned_ave):
  row_dots          (s_a, z_a) = audio . (pooling_proj, weight_proj);  (s_v, z_v) likewise   [:150-183]
  merge_weights2    masked softmax pooling over time per modality (own length arrays), 2-way
                    softmax -> acoustic_weight / visual_weight                               [:150-194]
  scale_add_rows    z = w_a * audio + w_v * video                                            [:197-199]
  ffn_fused         LN_final(W2 swish(W1 z + b1) + b2)  (no residual, no pre-norm)           [:197-208]
"""
from __future__ import annotations

import torch

from .. import engine, ops
from ..espnet_compat import LayerNorm, PositionwiseFeedForward
from .audiovisual_fusion_abs_module import AudioVisualFusionAbsModule


class AdaptiveAudioVisualFusion(AudioVisualFusionAbsModule):
    """Constructor mirrors adaptive_audiovisual_fusion.py:45-56."""

    def __init__(self, input_size: int, output_size: int = 256, hidden_units: int = 2048,
                 audiovisual_layer_type: str = "upsampling_positionwise",
                 merge_method: str = "learned_ave", activation_type: str = "swish",
                 acoustic_weight: float = 0.5, dropout_rate: float = 0.1,
                 acoustic_branch_drop_rate: float = 0.0):
        super().__init__()
        self.input_size = input_size
        self._output_size = output_size
        self.acoustic_weight = acoustic_weight
        self.acoustic_branch_drop_rate = acoustic_branch_drop_rate
        engine.act_code(activation_type)
        if audiovisual_layer_type != "upsampling_positionwise":
            raise ValueError("Support only upsampling positionwise feed forward fusion.")
        self.merge_method = merge_method
        if merge_method == "concat":
            # the reference builds PositionwiseFeedForward(idim=2*input_size) whose output (2*input)
            # then meets LayerNorm(output_size): it only runs when output_size == 2*input_size, and
            # no shipped config uses it
            raise NotImplementedError('merge_method="concat" of AdaptiveAudioVisualFusion is not '
                                      "built on the B200 path (no shipped config uses it)")
        elif merge_method == "learned_ave":
            self.acoustic_pooling_proj = torch.nn.Linear(input_size, 1)
            self.visual_pooling_proj = torch.nn.Linear(input_size, 1)
            self.acoustic_weight_proj = torch.nn.Linear(input_size, 1)
            self.visual_weight_proj = torch.nn.Linear(input_size, 1)
            self.audiovisual_layer = PositionwiseFeedForward(input_size, hidden_units, dropout_rate,
                                                             activation_type)
        elif merge_method == "fixed_ave":
            assert 0.0 <= acoustic_weight <= 1.0, "cgmlp weight should be between 0.0 and 1.0"
            self.audiovisual_layer = PositionwiseFeedForward(input_size, hidden_units, dropout_rate,
                                                             activation_type)
        else:
            raise ValueError(f"Unknow merge method: {merge_method}")
        self.norm_final = LayerNorm(output_size)
        self._packed = engine.PackedCache()

    def output_size(self) -> int:
        return self._output_size

    # ---------------------------------------------------------------------------------------
    def run(self, audio2d: torch.Tensor, video2d: torch.Tensor, lens_a: torch.Tensor,
            lens_v: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """Core on 2-D (B*T, d) activations; returns the fused (B*T, d) tensor and publishes
        acoustic_weight / visual_weight like the reference (:194)."""
        d = self.input_size
        dev = audio2d.device
        if d != 256 or self._output_size != 256:
            raise NotImplementedError("the B200 row-complete epilogue is built for size=256")
        if self.training and (self.acoustic_branch_drop_rate > 0 or self.audiovisual_layer.dropout_rate > 0):
            raise NotImplementedError("a no-grad call in train() mode with dropout / branch drop "
                                      "enabled: the inference kernels have no random paths; use "
                                      ".eval() (or a grad-mode call: training.py)")
        if self.merge_method == "learned_ave":
            ap, aw = self.acoustic_pooling_proj, self.acoustic_weight_proj
            vp, vw = self.visual_pooling_proj, self.visual_weight_proj
            vec = self._packed.get(
                "vec", [ap.weight, ap.bias, aw.weight, aw.bias, vp.weight, vp.bias, vw.weight, vw.bias],
                lambda: dict(pa=ap.weight.reshape(-1).contiguous(), wa=aw.weight.reshape(-1).contiguous(),
                             pv=vp.weight.reshape(-1).contiguous(), wv=vw.weight.reshape(-1).contiguous(),
                             sc=[float(ap.bias), float(vp.bias), float(aw.bias), float(vw.bias)]))
            d1, d2 = ops.row_dots(audio2d, vec["pa"], vec["wa"], video2d, vec["pv"], vec["wv"])
            sc = vec["sc"]
            w_a, w_v = ops.merge_weights2(d1, 1, d2, 1, lens_a, lens_v, sc[0], sc[1], sc[2], sc[3],
                                          d, B, T)
            self.acoustic_weight = w_a.view(B, 1, 1)
            self.visual_weight = w_v.view(B, 1, 1)
        else:
            w_a, w_v = self._packed.get(
                "fixedw" + str((B, str(dev), float(self.acoustic_weight))), [self.norm_final.weight],
                lambda: (torch.full((B,), float(self.acoustic_weight), device=dev, dtype=torch.float32),
                         torch.full((B,), 1.0 - float(self.acoustic_weight), device=dev,
                                    dtype=torch.float32)))
        # the weighted average only feeds the fusion FFN: operand storage
        z = ops.scale_add_rows(audio2d, video2d, w_a, w_v, T, out_dtype=engine.act_dtype())
        out = torch.empty((B * T, d), device=dev, dtype=torch.float32)
        # audiovisual_layer + norm_final: the FFN kernel without residual, norm_final as its LN0
        engine.ffn_block(None, z, self.audiovisual_layer, out_main=out, alpha=1.0,
                         ln0=(self.norm_final.weight, self.norm_final.bias), cache=self._packed,
                         key="avffn")
        return out

    def forward(self, audio_pad, audio_masks, video_pad, video_masks, cache=None):
        """Same contract as the reference forward (:113-131): returns (audiovisual (B,T,d),
        olens (B,))."""
        if cache is not None:
            raise NotImplementedError("cache is not None, which is not tested")
        engine.require_cuda(audio_pad, video_pad)
        if audio_pad.shape != video_pad.shape:
            raise NotImplementedError("the B200 fusion expects time-aligned streams of equal shape "
                                      "(avsr_espnet_model.py:439 aligns them)")
        B, T, d = audio_pad.shape
        a2 = audio_pad.reshape(B * T, d).contiguous().float()
        v2 = video_pad.reshape(B * T, d).contiguous().float()
        la = engine.lens_from_mask(audio_masks, B, T, a2.device)
        lv = engine.lens_from_mask(video_masks, B, T, a2.device)
        from .. import training
        if training.wants_grad(self, audio_pad, video_pad):
            out = training.fusion_forward(self, a2, v2, la, lv, B, T).view(B, T, d)
        else:
            out = self.run(a2, v2, la, lv, B, T).view(B, T, d)
        if audio_masks is None or video_masks is None:
            olens = torch.full((B,), T, dtype=torch.int64, device=a2.device)
        else:
            olens = torch.logical_or(audio_masks, video_masks).squeeze(1).sum(1)
        return out, olens
