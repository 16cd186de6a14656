"""TEST INFRASTRUCTURE: generate tests/golden/*.npz by running the REAL reference
(/root/reference/src/encoder/**, src/ctc/ctc.py, unmodified, over oracle/espnet_shim) on the seeded
cases of oracle/cases.py.  Run in the build container:  python -m oracle.gen_golden
The GPU box has no /root/reference; it checks against these stored outputs."""
from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, reference_loader, synth  # noqa: E402
from oracle.ref_path import make_valid_mask, rel_pos_emb  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def build_reference(ref, name):
    c = cases.CASES[name]
    cfg = copy.deepcopy(c["cfg"])
    if c["kind"] == "single":
        enc = ref.MyBranchformerEncoder(input_size=c["input_size"], **cfg)
    elif c["kind"] == "tailored":
        enc = ref.TailoredEncoder(embed_pos_enc_layer_type="rel_pos", embed_rel_pos_type="latest", **cfg)
    else:
        sub = {k: v for k, v in cfg.items() if k != "output_size"}  # passed explicitly (:74-83)
        a = dict(sub, encoder_class_type="branchformer")
        v = dict(sub, encoder_class_type="branchformer")
        enc = ref.ConventionalEncoder(input_size=256, acoustic_encoder_conf=a, visual_encoder_conf=v,
                                      output_size=cfg["output_size"], **c.get("wrap", {}))
    ctc = ref.CTC(odim=c["vocab"], encoder_output_size=cfg["output_size"], dropout_rate=0.0,
                  ctc_type="builtin", reduce=True)
    if cfg.get("interctc_use_conditioning", False) or c.get("wrap", {}).get("interctc_use_conditioning", False):
        enc.conditioning_layer = torch.nn.Linear(c["vocab"], cfg["output_size"])  # espnet_model.py:106-112
    enc.eval()
    ctc.eval()
    synth.fill_module(enc, seed=c["seed"], hot=c.get("hot", False))
    synth.fill_module(ctc, seed=c["seed"], prefix="ctc.", hot=c.get("hot", False))
    fk = cases.fusion_kwargs(name)
    if fk is not None:
        fusion = ref.AdaptiveAudioVisualFusion(**fk)
        fusion.eval()
        synth.fill_module(fusion, seed=c["seed"], prefix="fusion.")
        return enc, ctc, fusion
    return enc, ctc, None


def run_case(ref, name):
    c = cases.CASES[name]
    inp = cases.make_inputs(name)
    enc, ctc, fusion = build_reference(ref, name)
    out = {}
    with torch.no_grad():
        if c["kind"] == "single":
            y, olens, _ = enc(inp["x"], inp["lens"], ctc=ctc, max_layer=c.get("max_layer"))
            streams = {}
            if isinstance(y, tuple):
                y, inter = y
                for idx, t in inter:
                    streams[f"inter_{idx}"] = t
            streams["out"] = y
            weights = [(getattr(l, "weight_global", None), getattr(l, "weight_local", None))
                       for l in enc.encoders]
        else:
            d = c["cfg"]["output_size"]
            T = c["T"]
            pos = rel_pos_emb(T, d)
            mask = make_valid_mask(inp["lens"], T)
            mask_v = make_valid_mask(inp["lens_video"], T)
            ya, _, yv, _, _ = enc((inp["audio"], pos), mask, (inp["video"], pos), mask_v, ctc=ctc,
                                  audiovisual_fusion=fusion)
            streams = {}
            if isinstance(ya, tuple):
                ya, inter = ya
                for idx, t in inter:
                    streams[f"inter_{idx}"] = t
            streams.update({"out": ya, "out_video": yv})
            olens = inp["lens"]
            if fusion is not None:
                # avsr_espnet_model.py:467: the fused stream is what CTC sees
                fused, olens = fusion(ya, mask, yv, mask_v)
                streams["fused"] = fused
                aw = fusion.acoustic_weight
                out["acoustic_weight"] = (aw.flatten().numpy() if torch.is_tensor(aw)
                                          else np.array([aw], dtype=np.float32))
            if c["kind"] == "conventional":
                weights = [(getattr(l, "weight_global", None), getattr(l, "weight_local", None))
                           for l in enc.acoustic_encoder.encoders]
            else:
                weights = []
        y = streams.get("fused", streams["out"])
        tl = cases.target_lens(name, olens)
        loss = ctc(y, olens, inp["ys_pad"], tl)
        ctc.reduce = False
        loss_vec = ctc(y, olens, inp["ys_pad"], tl)
        amax = ctc.argmax(y)
        logp = ctc.log_softmax(y)
    st, sd_ = c.get("stride_t", 1), c.get("stride_d", 1)
    for k, v in streams.items():
        out[k] = v[:, ::st, ::sd_].numpy()
    out["olens"] = olens.numpy()
    out["tlens"] = tl.numpy()
    out["ctc_loss"] = np.array(loss.item(), dtype=np.float64)
    out["ctc_loss_vec"] = loss_vec.numpy()
    out["argmax"] = amax.numpy().astype(np.int16)
    out["logp_sample"] = logp[:, ::max(st, 4), :].numpy()
    wg = [w[0].flatten().numpy() for w in weights if w[0] is not None and torch.is_tensor(w[0])]
    if wg:
        out["weight_global"] = np.stack(wg)
    out["n_params"] = np.array(sum(p.numel() for p in enc.parameters()))
    return out


def main():
    ref = reference_loader.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]  # optional: regenerate just the named cases
    man = os.path.join(GOLDEN_DIR, "MANIFEST.json")
    meta = {"torch": torch.__version__, "cases": {}}
    if only and os.path.exists(man):
        with open(man) as f:
            meta = json.load(f)
    for name in (only or cases.CASES):
        res = run_case(ref, name)
        np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **res)
        meta["cases"][name] = {"ctc_loss": float(res["ctc_loss"]), "n_params": int(res["n_params"]),
                               "out_shape": list(res["out"].shape)}
        print(name, meta["cases"][name], flush=True)
    with open(man, "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
