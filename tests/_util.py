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
                                  output_size=cfg["output_size"])
    ctc = CTC(odim=c["vocab"], encoder_output_size=cfg["output_size"], dropout_rate=0.0,
              ctc_type="builtin", reduce=True)
    if cfg.get("interctc_use_conditioning", False):
        enc.conditioning_layer = torch.nn.Linear(c["vocab"], cfg["output_size"])
    enc.eval()
    ctc.eval()
    sd = synth.fill_module(enc, seed=c["seed"], hot=c.get("hot", False))
    sd.update(synth.fill_module(ctc, seed=c["seed"], prefix="ctc.", hot=c.get("hot", False)))
    return enc, ctc, sd


def run_oracle(name, sd):
    """CPU oracle on the seeded inputs.  Returns dict with out (B,T,d) [, out_video], olens, weights."""
    c = cases.CASES[name]
    inp = cases.make_inputs(name)
    res = {}
    with torch.no_grad():
        if c["kind"] == "single":
            y, olens, w = ref_path.branchformer_encoder(inp["x"], inp["lens"], sd, c["cfg"])
            res.update(out=y, olens=olens, weights=w)
            for idx, t in ref_path.branchformer_encoder.last_taps:
                res[f"inter_{idx}"] = t
        else:
            d = c["cfg"]["output_size"]
            T = c["T"]
            pos = ref_path.rel_pos_emb(T, d)
            mask = ref_path.make_valid_mask(inp["lens"], T)
            if c["kind"] == "tailored":
                a, v = ref_path.tailored_encoder(inp["audio"], pos, mask, inp["video"], pos, mask, sd,
                                                 c["cfg"])
                w = []
            else:
                a, v, w, _ = ref_path.conventional_encoder(inp["audio"], pos, mask, inp["video"], pos,
                                                           mask, sd, c["cfg"], c["cfg"])
            res.update(out=a, out_video=v, olens=inp["lens"], weights=w)
        tl = cases.target_lens(name, res["olens"])
        res["tlens"] = tl
        res["ctc_loss"] = ref_path.ctc_loss(res["out"], res["olens"], inp["ys_pad"], tl, sd, "ctc.ctc_lo")
        res["ctc_loss_vec"] = ref_path.ctc_loss(res["out"], res["olens"], inp["ys_pad"], tl, sd,
                                                "ctc.ctc_lo", reduce=False)
        res["argmax"] = torch.argmax(torch.nn.functional.linear(
            res["out"], sd["ctc.ctc_lo.weight"], sd["ctc.ctc_lo.bias"]), dim=2)
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
