"""Shared helpers of the parity tests: build the drop-in modules for a case of oracle/cases.py,
fill them with the seeded synthetic weights, run the CPU oracle on the same inputs."""
import copy
import os

import numpy as np
import torch

from oracle import cases, ref_path, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def build_dropin(name):
    """Returns (encoder, ctc) drop-in modules (on CPU) with synthetic weights, and the state dict."""
    from tailored_avsr_b200.ctc.ctc import CTC
    from tailored_avsr_b200.encoder.audiovisual.conventional.encoder import ConventionalEncoder
    from tailored_avsr_b200.encoder.audiovisual.tailored.encoder import TailoredEncoder
    from tailored_avsr_b200.encoder.branchformer.encoder import MyBranchformerEncoder
    c = cases.CASES[name]
    cfg = copy.deepcopy(c["cfg"])
    if c["kind"] == "single":
        enc = MyBranchformerEncoder(input_size=c["input_size"], **cfg)
    elif c["kind"] == "tailored":
        enc = TailoredEncoder(embed_pos_enc_layer_type="rel_pos", embed_rel_pos_type="latest", **cfg)
    else:
        sub = {k: v for k, v in cfg.items() if k != "output_size"}
        enc = ConventionalEncoder(input_size=256,
                                  acoustic_encoder_conf=dict(sub, encoder_class_type="branchformer"),
                                  visual_encoder_conf=dict(sub, encoder_class_type="branchformer"),
                                  output_size=cfg["output_size"], **c.get("wrap", {}))
    ctc = CTC(odim=c["vocab"], encoder_output_size=cfg["output_size"], dropout_rate=0.0,
              ctc_type="builtin", reduce=True)
    if cfg.get("interctc_use_conditioning", False) or c.get("wrap", {}).get("interctc_use_conditioning", False):
        enc.conditioning_layer = torch.nn.Linear(c["vocab"], cfg["output_size"])
    enc.eval()
    ctc.eval()
    sd = synth.fill_module(enc, seed=c["seed"], hot=c.get("hot", False))
    sd.update(synth.fill_module(ctc, seed=c["seed"], prefix="ctc.", hot=c.get("hot", False)))
    fusion = build_fusion(name)
    if fusion is not None:
        sd.update(synth.fill_module(fusion, seed=c["seed"], prefix="fusion."))
        enc.test_fusion = [fusion]  # in a list: not registered as a sub-module of the encoder
    return enc, ctc, sd


def build_fusion(name):
    """The case's AdaptiveAudioVisualFusion drop-in (CPU, weights not filled yet), or None."""
    from tailored_avsr_b200.audiovisual_fusion.adaptive_audiovisual_fusion import \
        AdaptiveAudioVisualFusion
    fk = cases.fusion_kwargs(name)
    if fk is None:
        return None
    return AdaptiveAudioVisualFusion(**fk).eval()


def run_oracle(name, sd):
    """CPU oracle on the seeded inputs.  Returns dict with out (B,T,d) [, out_video], olens, weights."""
    c = cases.CASES[name]
    inp = cases.make_inputs(name)
    res = {}
    with torch.no_grad():
        if c["kind"] == "single":
            y, olens, w = ref_path.branchformer_encoder(inp["x"], inp["lens"], sd, c["cfg"],
                                                        max_layer=c.get("max_layer"))
            res.update(out=y, olens=olens, weights=w)
            for idx, t in ref_path.branchformer_encoder.last_taps:
                res[f"inter_{idx}"] = t
        else:
            d = c["cfg"]["output_size"]
            T = c["T"]
            pos = ref_path.rel_pos_emb(T, d)
            mask = ref_path.make_valid_mask(inp["lens"], T)
            mask_v = ref_path.make_valid_mask(inp["lens_video"], T)
            fk = cases.fusion_kwargs(name)
            fuse = None
            if fk is not None:
                def fuse(a_, ma_, v_, mv_):
                    return ref_path.adaptive_av_fusion(
                        a_, ma_, v_, mv_, sd, "fusion.", merge_method=fk["merge_method"],
                        acoustic_weight=fk["acoustic_weight"], act=fk["activation_type"])
            if c["kind"] == "tailored":
                r = ref_path.tailored_encoder(
                    inp["audio"], pos, mask, inp["video"], pos, mask_v, sd, c["cfg"],
                    fusion=(lambda *a_: fuse(*a_)[0]) if fuse else None,
                    ctc_softmax=lambda h: ref_path.ctc_log_softmax(h, sd, "ctc.ctc_lo").exp())
                a, v = r[0], r[1]
                for idx, t in (r[2] if len(r) > 2 else []):
                    res[f"inter_{idx}"] = t
                w = []
            elif c.get("wrap", {}).get("interctc_layer_idx"):
                a, v, taps = ref_path.conventional_encoder_interctc(
                    inp["audio"], pos, mask, inp["video"], pos, mask_v, sd, c["cfg"], c["wrap"],
                    fusion=lambda *a_: fuse(*a_)[0],
                    ctc_softmax=lambda h: ref_path.ctc_log_softmax(h, sd, "ctc.ctc_lo").exp())
                for idx, t in taps:
                    res[f"inter_{idx}"] = t
                w = []
            else:
                a, v, w, _ = ref_path.conventional_encoder(inp["audio"], pos, mask, inp["video"], pos,
                                                           mask_v, sd, c["cfg"], c["cfg"])
            res.update(out=a, out_video=v, olens=inp["lens"], weights=w)
            res["lens_video"] = inp["lens_video"]
            if fuse is not None:
                fused, folens, aw = fuse(a, mask, v, mask_v)
                res.update(fused=fused, olens=folens, lens_audio=inp["lens"], acoustic_weight=aw)
        # the CTC input: the fused stream when a fusion module sits behind the encoder
        res["hs"] = res.get("fused", res["out"])
        tl = cases.target_lens(name, res["olens"])
        res["tlens"] = tl
        res["ctc_loss"] = ref_path.ctc_loss(res["hs"], res["olens"], inp["ys_pad"], tl, sd, "ctc.ctc_lo")
        res["ctc_loss_vec"] = ref_path.ctc_loss(res["hs"], res["olens"], inp["ys_pad"], tl, sd,
                                                "ctc.ctc_lo", reduce=False)
        res["argmax"] = torch.argmax(torch.nn.functional.linear(
            res["hs"], sd["ctc.ctc_lo.weight"], sd["ctc.ctc_lo.bias"]), dim=2)
    res["inputs"] = inp
    return res


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))


def valid_frames(t, lens):
    """Concatenate the valid frames (t < lens[b]) of a (B,T,...) tensor."""
    return torch.cat([t[b, : int(lens[b])] for b in range(t.shape[0])], 0)


def rel_errors(y, ref, lens):
    """(max abs / max abs, Frobenius) over valid frames — the parity metrics of SURVEY.md §8d."""
    a = valid_frames(y.double().cpu(), lens)
    b = valid_frames(ref.double().cpu(), lens)
    return (float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm()))
