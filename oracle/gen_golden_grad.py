"""TEST INFRASTRUCTURE: golden GRADIENTS of the hot path for the training rows of SURVEY.md §8: the
REAL reference modules run live over oracle/espnet_shim with autograd on, loss = CTC loss of the
case (on the fused stream for the audio-visual cases); per parameter and input the gradient's L2
norm, its sum and a strided sample are stored.  Two sets: eval mode (dropout and stochastic depth
are identities, the arithmetic is the training arithmetic) and train() mode with every dropout
active and masks injected from oracle/dropmask.py (grad_<case>_dropout.npz).  The functional port
(oracle/ref_path.py) must reproduce the eval set (tests/test_oracle_cpu.py), and the CUDA training
path both (tests/test_backward_gpu.py).  Run in the build container:  python -m oracle.gen_golden_grad"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, dropmask, gen_golden, reference_loader  # noqa: E402

CASES = ["vsr_small", "asr_small", "asr_tailored_small", "vsr_tailored_small", "concat_small",
         "asr_interctc_cond"]
INTERCTC_WEIGHT = 0.5   # InterCTC cases: loss = CTC(final) + 0.5 * sum over taps of CTC(tap)
# train() mode with every dropout site active (rates of the case's config, 0.1 like the shipped
# YAMLs), masks from oracle/dropmask.py::MaskSource(GOLDEN_SEED): grad_<case>_dropout.npz
DROPOUT_CASES = ["vsr_small", "asr_small", "vsr_tailored_small", "concat_small",
                 "av_fusion_conventional", "av_fusion_tailored", "asr_interctc_cond",
                 "av_tailored_interctc", "av_conventional_interctc"]
# audio-visual: ConventionalEncoder + AdaptiveAudioVisualFusion + CTC on the fused stream
# (avsr_espnet_model.py:467,678), different audio / video masks
AV_CASES = ["av_fusion_conventional", "av_fusion_tailored", "av_tailored_interctc",
            "av_tailored_interctc_sep", "av_conventional_interctc"]


def summarize(named_grads):
    out = {}
    for name, g in named_grads:
        g = g.detach().double().reshape(-1)
        out["norm/" + name] = np.array(float(g.norm()))
        out["sum/" + name] = np.array(float(g.sum()))
        out["sample/" + name] = g[:: max(1, g.numel() // 16)][:16].numpy()
    return out


def main():
    ref = reference_loader.load()
    for name, drop in [(n, False) for n in CASES + AV_CASES] + [(n, True) for n in DROPOUT_CASES]:
        c = cases.CASES[name]
        inp = cases.make_inputs(name)
        enc, ctc, fusion = gen_golden.build_reference(ref, name)
        src = dropmask.MaskSource(dropmask.GOLDEN_SEED)
        if drop:
            enc.train()
            if fusion is not None:
                fusion.train()
        if c["kind"] == "single":
            x = inp["x"].clone().requires_grad_(True)
            with dropmask.patched_dropout(src):
                y, olens, _ = enc(x, inp["lens"], ctc=ctc)
                tl = cases.target_lens(name, olens)
                taps = []
                if isinstance(y, tuple):
                    y, taps = y
                loss = ctc(y, olens, inp["ys_pad"], tl)
                for _, tap in taps:
                    loss = loss + INTERCTC_WEIGHT * ctc(tap, olens, inp["ys_pad"], tl)
            inputs = [("input", x)]
        else:
            from oracle.ref_path import make_valid_mask, rel_pos_emb
            d, T = c["cfg"]["output_size"], c["T"]
            pos = rel_pos_emb(T, d)
            mask, mask_v = make_valid_mask(inp["lens"], T), make_valid_mask(inp["lens_video"], T)
            a = inp["audio"].clone().requires_grad_(True)
            v = inp["video"].clone().requires_grad_(True)
            with dropmask.patched_dropout(src):
                ya, _, yv, _, _ = enc((a, pos), mask, (v, pos), mask_v, ctc=ctc, audiovisual_fusion=fusion)
                taps = []
                if isinstance(ya, tuple):
                    ya, taps = ya
                y, olens = fusion(ya, mask, yv, mask_v)
                tl = cases.target_lens(name, olens)
                loss = ctc(y, olens, inp["ys_pad"], tl)
                for _, tap in taps:      # the fused intermediate outputs (avsr_espnet_model.py)
                    loss = loss + INTERCTC_WEIGHT * ctc(tap, olens, inp["ys_pad"], tl)
            inputs = [("input_audio", a), ("input_video", v)]
        loss.backward()
        if drop:
            print(name, "dropout sites:", len(src.calls), src.calls[:14])
            name = name + "_dropout"
        grads = [("enc." + n, p.grad) for n, p in enc.named_parameters() if p.grad is not None]
        grads += [("ctc." + n, p.grad) for n, p in ctc.named_parameters()]
        if fusion is not None:
            grads += [("fusion." + n, p.grad) for n, p in fusion.named_parameters() if p.grad is not None]
        grads += [(n, t.grad) for n, t in inputs]
        out = summarize(grads)
        out["loss"] = np.array(float(loss))
        if drop:
            out["n_masks"] = np.array(len(src.calls))
            out["out_sample"] = y.detach().double().reshape(-1)[::97][:64].numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"grad_{name}.npz"), **out)
        print(name, float(loss), len(grads), "gradients")


if __name__ == "__main__":
    main()
