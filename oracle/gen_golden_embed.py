"""TEST INFRASTRUCTURE: golden vectors of the AVSR embedding path, produced by the REAL reference
(src/embedding_for_avsr/default.py run live over oracle/espnet_shim, plus the model's
audiovisual_alignment arithmetic restated from src/models/avsr_espnet_model.py:512-541, which
cannot be imported here: the model file pulls in espnet2's task machinery).
Run in the build container:  python -m oracle.gen_golden_embed"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_path, reference_loader, synth  # noqa: E402

CASE = dict(B=3, Ta=203, Fa=80, Tv=52, Fv=512, lens_a=[203, 150, 77], lens_v=[52, 38, 19], d=256, seed=41)


def inputs():
    c = CASE
    return (synth.randn((c["B"], c["Ta"], c["Fa"]), c["seed"]), torch.tensor(c["lens_a"]),
            synth.randn((c["B"], c["Tv"], c["Fv"]), c["seed"] + 1), torch.tensor(c["lens_v"]))


def main():
    reference_loader.load()
    from src.embedding_for_avsr.default import DefaultEmbeddingLayerForAVSR
    c = CASE
    ae = DefaultEmbeddingLayerForAVSR(c["Fa"], c["d"], input_layer="conv2d").eval()
    ve = DefaultEmbeddingLayerForAVSR(c["Fv"], c["d"], input_layer="linear").eval()
    synth.fill_module(ae, seed=c["seed"], prefix="acoustic_embed.")
    synth.fill_module(ve, seed=c["seed"], prefix="visual_embed.")
    xa, la, xv, lv = inputs()
    with torch.no_grad():
        a, ma = ae.apply_embed_layer(xa, la)
        v, mv = ve.apply_embed_layer(xv, lv)
        a2, ma2, v2, mv2 = ref_path.audiovisual_alignment(a, ma, v, mv)
        (ap, pos_a) = ae.apply_pos_enc(a2)
        (vp, pos_v) = ve.apply_pos_enc(v2)
        (fa, fpos), fm = ae(xa, la)
    out = dict(audio_embed=a.numpy(), audio_mask=ma.numpy(), video_embed=v.numpy(), video_mask=mv.numpy(),
               audio_in=ap.numpy(), video_in=vp.numpy(), pos=pos_a.numpy()[:, ::4],
               audio_mask_aligned=ma2.numpy(), video_mask_aligned=mv2.numpy(),
               forward_audio=fa.numpy()[:, ::2, ::2])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "avsr_embed.npz"), **out)
    print({k: v_.shape for k, v_ in out.items()})


if __name__ == "__main__":
    main()
