"""B200 drop-in for the reference `TailoredEncoderLayer`
(src/encoder/audiovisual/tailored/encoder_layer.py:50-274): FFN-macaron / FFN / norm_final are
shared by the two streams, each stream owns ONE branch (rel-pos MHSA or cgMLP) with its own
residual.  Both streams go through the shared FFN weights in a single GEMM launch (2*B*T rows)."""
from __future__ import annotations

from typing import Optional

import torch

from .... import engine, ops
from ....espnet_compat import LayerNorm


class TailoredEncoderLayer(torch.nn.Module):
    def __init__(self, size: int, feed_forward_macaron: Optional[torch.nn.Module],
                 acoustic_attn: Optional[torch.nn.Module], acoustic_cgmlp: Optional[torch.nn.Module],
                 visual_attn: Optional[torch.nn.Module], visual_cgmlp: Optional[torch.nn.Module],
                 feed_forward: Optional[torch.nn.Module], dropout_rate: float,
                 acoustic_branch_drop_rate: float = 0.0, stochastic_depth_rate: float = 0.0):
        super().__init__()
        self.size = size
        self.ff_scale = 0.5
        self.feed_forward_macaron = feed_forward_macaron
        self.norm_ff_macaron = LayerNorm(size)
        self.acoustic_attn = acoustic_attn
        if self.acoustic_attn is not None:
            self.acoustic_norm_mha = LayerNorm(size)
        self.acoustic_cgmlp = acoustic_cgmlp
        if self.acoustic_cgmlp is not None:
            self.acoustic_norm_cgmlp = LayerNorm(size)
        self.visual_attn = visual_attn
        if self.visual_attn is not None:
            self.visual_norm_mha = LayerNorm(size)
        self.visual_cgmlp = visual_cgmlp
        if self.visual_cgmlp is not None:
            self.visual_norm_cgmlp = LayerNorm(size)
        self.feed_forward = feed_forward
        self.norm_ff = LayerNorm(size)
        self.norm_final = LayerNorm(size)
        self.dropout = torch.nn.Dropout(dropout_rate)
        self.acoustic_branch_drop_rate = acoustic_branch_drop_rate
        self.stochastic_depth_rate = stochastic_depth_rate
        self._packed = engine.PackedCache()

    def _check_supported(self):
        if self.size != 256:
            raise NotImplementedError("the B200 row-complete GEMM epilogue is built for size=256")
        if self.feed_forward_macaron is None or self.feed_forward is None:
            raise NotImplementedError("macaron=False is not built on the B200 path")
        for tag in ("acoustic", "visual"):
            a, c = getattr(self, tag + "_attn"), getattr(self, tag + "_cgmlp")
            if (a is not None) and (c is not None):
                raise RuntimeError(f"Only one of the possible {tag} tailored modules should be not "
                                   f"None: {a}, {c}.")
            if a is None and c is None:
                raise NotImplementedError(f"{tag} stream without a tailored module is not built")
            if a is not None and (a.d_k != 64 or a.h * a.d_k != self.size):
                raise NotImplementedError(
                    f"the B200 attention kernel is built for head width d_k = 64; {tag} attention "
                    f"has h={a.h}, d_k={a.d_k}")
        if self.training and not torch.is_grad_enabled() and (
                self.dropout.p > 0 or self.stochastic_depth_rate > 0):
            raise NotImplementedError("a no-grad call in train() mode with dropout / stochastic depth "
                                      "enabled: the inference kernels have no random paths; use "
                                      ".eval() (or a grad-mode call: training.py)")

    def run(self, x, xn, pos_proj_a, pos_proj_v, lens_a, lens_v, B, T, next_norm=None,
            next_norm_dtype=None):
        """Core on 2-D activations holding BOTH streams stacked: rows [0, B*T) audio,
        [B*T, 2*B*T) video.  x: block input (fp32), xn: norm_ff_macaron(x) in operand storage.
        Returns (y, yn); yn is stored as `next_norm_dtype` (default: operand storage)."""
        d = self.size
        M = B * T
        dev = x.device
        adt = engine.act_dtype()
        new2 = lambda dt=torch.float32: torch.empty((2 * M, d), device=dev, dtype=dt)  # noqa: E731
        # shared macaron FFN over both streams; per-stream branch norms differ -> plain output, then
        # the stream-specific LayerNorm is fused as lnA of two half-height launches below.
        x_a = new2()
        xb_in = new2(adt)  # per-stream branch input (its own LayerNorm)
        for s, tag in enumerate(("acoustic", "visual")):
            norm = getattr(self, f"{tag}_norm_mha", None) if getattr(self, f"{tag}_attn") is not None \
                else getattr(self, f"{tag}_norm_cgmlp")
            sl = slice(s * M, (s + 1) * M)
            engine.ffn_block(x[sl], xn[sl], self.feed_forward_macaron, out_main=x_a[sl],
                             lnA=(norm.weight, norm.bias), out_lnA=xb_in[sl], cache=self._packed,
                             key="ffm")
        x_b = new2()
        xf = new2(adt)
        lnF = (self.norm_ff.weight, self.norm_ff.bias)
        for s, (tag, pos_proj, lens) in enumerate((("acoustic", pos_proj_a, lens_a),
                                                   ("visual", pos_proj_v, lens_v))):
            sl = slice(s * M, (s + 1) * M)
            attn = getattr(self, f"{tag}_attn")
            if attn is not None:
                if pos_proj is None:
                    raise NotImplementedError("attention without relative positional embedding is "
                                              "not built on the B200 path")
                ctx = engine.attention_ctx(xb_in[sl], attn, pos_proj, lens, B, T, self._packed,
                                           f"{tag}_qkv")
                engine.linear_rowln(ctx, attn.linear_out.weight, attn.linear_out.bias, self._packed,
                                    f"{tag}_lo", residual=x_a[sl], alpha=1.0, out_main=x_b[sl],
                                    lnA=lnF, out_lnA=xf[sl])
            else:
                cg = getattr(self, f"{tag}_cgmlp")
                u = engine.cgmlp_gated(xb_in[sl], cg, B, T, self._packed, f"{tag}_conv")
                engine.linear_rowln(u, cg.channel_proj2.weight, cg.channel_proj2.bias, self._packed,
                                    f"{tag}_p2", residual=x_a[sl], alpha=1.0, out_main=x_b[sl],
                                    lnA=lnF, out_lnA=xf[sl])
        y = new2()
        yn = new2(next_norm_dtype or adt) if next_norm is not None else None
        engine.ffn_block(x_b, xf, self.feed_forward, out_main=y,
                         ln0=(self.norm_final.weight, self.norm_final.bias),
                         lnA=next_norm, out_lnA=yn, cache=self._packed, key="ff")
        return y, yn

    def forward(self, audio_input, audio_masks, video_input, video_masks, cache=None):
        """Same contract as the reference forward (tailored/encoder_layer.py:118-137)."""
        if cache is not None:
            raise NotImplementedError("cache is not None, which is not tested")
        audio, audio_pos = audio_input if isinstance(audio_input, tuple) else (audio_input, None)
        video, video_pos = video_input if isinstance(video_input, tuple) else (video_input, None)
        self._check_supported()
        engine.require_cuda(audio, video)
        if audio.shape != video.shape:
            raise NotImplementedError("the B200 tailored layer expects time-aligned streams of "
                                      "equal shape (avsr_espnet_model.py:439 aligns them)")
        B, T, d = audio.shape
        M = B * T
        from .... import training
        if training.wants_grad(self, audio, video):
            from .... import ops_backward as ob
            dev = audio.device

            def pair(pos):
                if pos is None:
                    return None
                p2 = pos.reshape(-1, d).contiguous().float()
                return p2, ob.transpose_2d(p2, pad=True)

            a_out, v_out = training.tailored_layer_forward(
                self, audio.reshape(M, d).contiguous().float(), video.reshape(M, d).contiguous().float(),
                B, T, engine.lens_from_mask(audio_masks, B, T, dev),
                engine.lens_from_mask(video_masks, B, T, dev), pair(audio_pos), pair(video_pos))
            a_out, v_out = a_out.view(B, T, d), v_out.view(B, T, d)
            a_ret = (a_out, audio_pos) if audio_pos is not None else a_out
            v_ret = (v_out, video_pos) if video_pos is not None else v_out
            return a_ret, audio_masks, v_ret, video_masks
        x = torch.cat([audio.reshape(M, d), video.reshape(M, d)], 0).contiguous().float()
        xn = ops.layernorm(x, self.norm_ff_macaron.weight, self.norm_ff_macaron.bias, eps=1e-12,
                           out_dtype=engine.act_dtype())
        la = engine.lens_from_mask(audio_masks, B, T, x.device)
        lv = engine.lens_from_mask(video_masks, B, T, x.device)
        pa = pv = None
        if self.acoustic_attn is not None and audio_pos is not None:
            pa = engine.pos_projection(self.acoustic_attn, audio_pos.float(), self._packed)
        if self.visual_attn is not None and video_pos is not None:
            pv = engine.pos_projection(self.visual_attn, video_pos.float(), self._packed)
        y, _ = self.run(x, xn, pa, pv, la, lv, B, T)
        a_out, v_out = y[:M].view(B, T, d), y[M:].view(B, T, d)
        a_ret = (a_out, audio_pos) if audio_pos is not None else a_out
        v_ret = (v_out, video_pos) if video_pos is not None else v_out
        return a_ret, audio_masks, v_ret, video_masks
