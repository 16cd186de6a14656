"""B200 drop-in for the reference `ConventionalEncoder`
(src/encoder/audiovisual/conventional/encoder.py:35-225): two independent Branchformer stacks
(acoustic, visual) that share nothing; they are issued on two CUDA streams so the GPU can overlap
them (the reference runs them back to back, :149-150)."""
from __future__ import annotations

from typing import List

import torch

from .... import engine
from ...branchformer.encoder import MyBranchformerEncoder
from ..audiovisual_abs_encoder import AudioVisualAbsEncoder


class ConventionalEncoder(AudioVisualAbsEncoder):
    """Constructor mirrors conventional/encoder.py:39-50."""

    def __init__(self, input_size, acoustic_encoder_conf, visual_encoder_conf, output_size: int = 256,
                 embed_pos_enc_layer_type: str = "rel_pos", embed_rel_pos_type: str = "latest",
                 interctc_use_conditioning: bool = False,
                 audiovisual_interctc_conditioning: bool = False,
                 interctc_layer_idx: List[int] = []):
        super().__init__()
        assert (embed_pos_enc_layer_type == acoustic_encoder_conf["pos_enc_layer_type"]
                == visual_encoder_conf["pos_enc_layer_type"]), (
            embed_pos_enc_layer_type, acoustic_encoder_conf["pos_enc_layer_type"],
            visual_encoder_conf["pos_enc_layer_type"])
        assert (embed_rel_pos_type == acoustic_encoder_conf["rel_pos_type"]
                == visual_encoder_conf["rel_pos_type"]), (
            embed_rel_pos_type, acoustic_encoder_conf["rel_pos_type"], visual_encoder_conf["rel_pos_type"])
        acoustic_cls = self.get_encoder_class(acoustic_encoder_conf["encoder_class_type"])
        visual_cls = self.get_encoder_class(visual_encoder_conf["encoder_class_type"])
        # the reference deletes the key from the caller's dicts in place (:71-72); kept for parity
        del acoustic_encoder_conf["encoder_class_type"]
        del visual_encoder_conf["encoder_class_type"]
        self.acoustic_encoder = acoustic_cls(input_size=input_size, output_size=output_size,
                                             **acoustic_encoder_conf)
        self.visual_encoder = visual_cls(input_size=input_size, output_size=output_size,
                                         **visual_encoder_conf)
        assert len(self.acoustic_encoder.encoders) == len(self.visual_encoder.encoders), \
            "Both encoders must have the same number of blocks."
        assert self.acoustic_encoder.output_size() == self.visual_encoder.output_size(), \
            "Output size should be the same in both wrapped encoders."
        assert self.acoustic_encoder.embed is None and self.visual_encoder.embed is None, \
            "The embedding layers of both encoders should be None."
        assert (len(self.acoustic_encoder.interctc_layer_idx) == 0
                and len(self.visual_encoder.interctc_layer_idx) == 0), \
            "InterCTC loss must be defined in the WrapperEncoder."
        assert (self.acoustic_encoder.interctc_use_conditioning is False
                and self.visual_encoder.interctc_use_conditioning is False), \
            "InterCTC conditioning must be defined in the WrapperEncoder."
        num_blocks = len(self.acoustic_encoder.encoders)
        self.interctc_layer_idx = interctc_layer_idx
        if len(interctc_layer_idx) > 0:
            assert 0 < min(interctc_layer_idx) and max(interctc_layer_idx) < num_blocks
        self.interctc_use_conditioning = interctc_use_conditioning
        self.audiovisual_interctc_conditioning = audiovisual_interctc_conditioning
        assert not (self.interctc_use_conditioning is False and self.audiovisual_interctc_conditioning is True), \
            "Audio-Visual InterCTC conditioning only can be applied if interctc_use_conditioning is set to True."
        self.conditioning_layer = None
        self._side_stream = None

    def output_size(self) -> int:
        return self.acoustic_encoder.output_size()

    def get_encoder_class(self, encoder_class_type):
        if encoder_class_type == "branchformer":
            return MyBranchformerEncoder
        if encoder_class_type == "conformer":
            raise NotImplementedError("encoder_class_type='conformer' is not built on the B200 path")
        raise ValueError("unknown encoder_class_type: " + encoder_class_type)

    def _run_stack(self, enc: MyBranchformerEncoder, x_pad, masks):
        xs, pos_emb = x_pad if isinstance(x_pad, tuple) else (x_pad, None)
        if pos_emb is None:
            raise NotImplementedError("the conventional AV encoder expects (x, pos_emb) inputs "
                                      "(avsr_espnet_model.py:447-451)")
        B, T, d = xs.shape
        first = enc.encoders[0]
        x, xn, pos_emb, masks, B, T = enc._embed((xs, pos_emb), masks,
                                                 (first.norm_ff_macaron.weight, first.norm_ff_macaron.bias))
        lens = engine.lens_from_mask(masks, B, T, x.device)
        out, _ = enc.run_blocks(x, xn, pos_emb, lens, B, T)
        return out.view(B, T, d)

    def forward(self, audio_pad, audio_masks, video_pad, video_masks, prev_states=None, ctc=None,
                audiovisual_fusion=None):
        """Same contract as the reference forward (conventional/encoder.py:116-144)."""
        a0 = audio_pad[0] if isinstance(audio_pad, tuple) else audio_pad
        v0 = video_pad[0] if isinstance(video_pad, tuple) else video_pad
        engine.require_cuda(a0, v0)
        from .... import training
        if training.wants_grad(self, a0, v0):
            # training step: the two stacks as autograd graphs of per-block nodes (training.py)
            return training.conventional_encoder_forward(self, audio_pad, audio_masks, video_pad,
                                                         video_masks, ctc=ctc, fusion=audiovisual_fusion)
        if len(self.interctc_layer_idx) > 0:
            return self._forward_interctc(audio_pad, audio_masks, video_pad, video_masks, ctc,
                                          audiovisual_fusion)
        cur = torch.cuda.current_stream(a0.device)
        if self._side_stream is None or self._side_stream.device != a0.device:
            self._side_stream = torch.cuda.Stream(device=a0.device)
        side = self._side_stream
        side.wait_stream(cur)
        audio_out = self._run_stack(self.acoustic_encoder, audio_pad, audio_masks)
        with torch.cuda.stream(side):
            video_out = self._run_stack(self.visual_encoder, video_pad, video_masks)
            video_out.record_stream(cur)
        cur.wait_stream(side)
        return audio_out, audio_masks, video_out, video_masks, None

    def _forward_interctc(self, audio_pad, audio_masks, video_pad, video_masks, ctc, fusion):
        """Layer-zipped form with audio-visual InterCTC taps (conventional/encoder.py:152-199): after
        the tapped blocks both streams are normalised, fused, and (with conditioning) get the CTC
        posteriors added back through `conditioning_layer`."""
        from .... import ops
        if fusion is None or not hasattr(fusion, "run"):
            raise ValueError("audio-visual InterCTC taps need the B200 `audiovisual_fusion` module "
                             "(conventional/encoder.py:171-177)")
        if self.interctc_use_conditioning and (ctc is None or self.conditioning_layer is None):
            raise ValueError("InterCTC self-conditioning needs the `ctc` module and an assigned "
                             "`conditioning_layer`")
        st = []
        for enc, x_pad, masks in ((self.acoustic_encoder, audio_pad, audio_masks),
                                  (self.visual_encoder, video_pad, video_masks)):
            xs, pos_emb = x_pad if isinstance(x_pad, tuple) else (x_pad, None)
            if pos_emb is None:
                raise NotImplementedError("the conventional AV encoder expects (x, pos_emb) inputs")
            first = enc.encoders[0]
            x, xn, pos_emb, masks, B, T = enc._embed(
                (xs, pos_emb), masks, (first.norm_ff_macaron.weight, first.norm_ff_macaron.bias))
            st.append(dict(enc=enc, x=x, xn=xn, pos=enc._pos_proj_all(pos_emb),
                           lens=engine.lens_from_mask(masks, B, T, x.device)))
        d = self.output_size()
        n = len(self.acoustic_encoder.encoders)
        inter = []
        for i in range(n):
            for s_ in st:
                enc = s_["enc"]
                layer = enc.encoders[i]
                layer._check_supported()
                if i + 1 < n:
                    nxt = enc.encoders[i + 1]
                    s_["next_norm"] = (nxt.norm_ff_macaron.weight, nxt.norm_ff_macaron.bias)
                else:
                    s_["next_norm"] = ((enc.after_norm.weight, enc.after_norm.bias)
                                       if enc.normalize_before else None)
                s_["nn_dtype"] = torch.float32 if i + 1 == n else None   # after_norm: fp32 output
                s_["x"], s_["xn"] = layer.run(s_["x"], s_["xn"], s_["pos"].get(i), s_["lens"], B, T,
                                              next_norm=s_["next_norm"], next_norm_dtype=s_["nn_dtype"])
            if (i + 1) in self.interctc_layer_idx:
                taps = []
                for s_ in st:
                    enc = s_["enc"]
                    taps.append(ops.layernorm(s_["x"], enc.after_norm.weight, enc.after_norm.bias, eps=1e-12)
                                if enc.normalize_before else s_["x"])
                fused = fusion.run(taps[0], taps[1], st[0]["lens"], st[1]["lens"], B, T)
                inter.append((i + 1, fused.view(B, T, d)))
                if self.interctc_use_conditioning:
                    cl = self.conditioning_layer
                    for s_, tap in zip(st, taps):
                        src = fused if self.audiovisual_interctc_conditioning else tap
                        prob = ctc.softmax(src.view(B, T, d)).reshape(B * T, -1).contiguous()
                        s_["x"], s_["xn"] = ops.vocab_residual(
                            s_["x"], prob, cl.weight.contiguous(), cl.bias, ln=s_["next_norm"],
                            ln_dtype=s_["nn_dtype"] or engine.act_dtype())
        outs = [(s_["xn"] if s_["enc"].normalize_before else s_["x"]).view(B, T, d) for s_ in st]
        return (outs[0], inter), audio_masks, outs[1], video_masks, None
