"""AVSR embedding layers: DefaultEmbeddingLayerForAVSR drop-in against the
CPU oracle port and the golden vectors of the live reference (oracle/gen_golden_embed.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden_embed, ref_path, reference_loader, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "avsr_embed.npz")


def _build():
    from tailored_avsr_b200.embedding_for_avsr.default import DefaultEmbeddingLayerForAVSR
    c = gen_golden_embed.CASE
    ae = DefaultEmbeddingLayerForAVSR(c["Fa"], c["d"], input_layer="conv2d").eval()
    ve = DefaultEmbeddingLayerForAVSR(c["Fv"], c["d"], input_layer="linear").eval()
    sd = synth.fill_module(ae, seed=c["seed"], prefix="acoustic_embed.")
    sd.update(synth.fill_module(ve, seed=c["seed"], prefix="visual_embed."))
    return ae, ve, sd


def _oracle(sd):
    xa, la, xv, lv = gen_golden_embed.inputs()
    a, ma = ref_path.avsr_embed_layer(xa, la, sd, "acoustic_embed.", "conv2d")
    v, mv = ref_path.avsr_embed_layer(xv, lv, sd, "visual_embed.", "linear")
    a2, ma2, v2, mv2 = ref_path.audiovisual_alignment(a, ma, v, mv)
    (ap, pos), (vp, _) = ref_path.rel_pos_enc(a2), ref_path.rel_pos_enc(v2)
    return dict(audio_embed=a, audio_mask=ma, video_embed=v, video_mask=mv, audio_in=ap, video_in=vp,
                pos=pos, audio_mask_aligned=ma2, video_mask_aligned=mv2)


def test_oracle_embed_port_matches_reference_golden():
    _, _, sd = _build()
    got, gold = _oracle(sd), np.load(GOLD)
    for k in ("audio_embed", "video_embed", "audio_in", "video_in"):
        assert np.allclose(got[k].numpy(), gold[k], rtol=1e-4, atol=1e-5), k
    for k in ("audio_mask", "video_mask", "audio_mask_aligned", "video_mask_aligned"):
        assert np.array_equal(got[k].numpy(), gold[k]), k
    assert np.allclose(got["pos"].numpy()[:, ::4], gold["pos"], atol=1e-6)


@pytest.mark.skipif(not reference_loader.available(), reason="/root/reference not present")
def test_embed_state_dict_layout_equals_reference():
    reference_loader.load()
    from src.embedding_for_avsr.default import DefaultEmbeddingLayerForAVSR as Ref
    ae, ve, _ = _build()
    c = gen_golden_embed.CASE
    for mine, theirs in ((ae, Ref(c["Fa"], c["d"], input_layer="conv2d")),
                         (ve, Ref(c["Fv"], c["d"], input_layer="linear"))):
        a = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in theirs.state_dict().items()}
        assert a == b
        theirs.load_state_dict(mine.state_dict(), strict=True)
        mine.load_state_dict(theirs.state_dict(), strict=True)
        assert mine.output_size() == theirs.output_size()


@pytest.mark.gpu
def test_embed_dropin_matches_oracle_and_golden_on_gpu():
    ae, ve, sd = _build()
    want, gold = _oracle(sd), np.load(GOLD)
    xa, la, xv, lv = gen_golden_embed.inputs()
    ae, ve = ae.cuda(), ve.cuda()
    with torch.no_grad():
        a, ma = ae.apply_embed_layer(xa.cuda(), la.cuda())
        v, mv = ve.apply_embed_layer(xv.cuda(), lv.cuda())
        a2, ma2, v2, mv2 = ref_path.audiovisual_alignment(a, ma, v, mv)   # torch F.pad: data movement
        (ap, pos), (vp, _) = ae.apply_pos_enc(a2), ve.apply_pos_enc(v2)
        (fa, fpos), fm = ae(xa.cuda(), la.cuda())

    def rel(x, y):
        x, y = x.double().cpu(), torch.as_tensor(y).double()
        return float((x - y).abs().max() / y.abs().max())

    assert rel(a, want["audio_embed"]) <= 1e-3 and rel(v, want["video_embed"]) <= 1e-3
    assert rel(ap, want["audio_in"]) <= 1e-3 and rel(vp, want["video_in"]) <= 1e-3
    assert rel(ap, gold["audio_in"]) <= 1e-3 and rel(vp, gold["video_in"]) <= 1e-3
    assert rel(fa[:, ::2, ::2], gold["forward_audio"]) <= 1e-3
    assert torch.equal(ma.cpu(), want["audio_mask"]) and torch.equal(mv.cpu(), want["video_mask"])
    assert torch.equal(ma2.cpu(), want["audio_mask_aligned"]) and torch.equal(fm.cpu(), want["audio_mask"])
    assert float((pos.cpu() - want["pos"]).abs().max()) < 1e-6 and fpos.shape == (1, 2 * 50 - 1, 256)


def test_embed_refuses_cpu_tensors():
    ae, _, _ = _build()
    with pytest.raises(Exception):
        ae.apply_embed_layer(torch.zeros(1, 23, 80), torch.tensor([23]))


@pytest.mark.gpu
def test_av_pipeline_from_raw_features_matches_oracle_composition():
    """The whole AVSR encode() call stack on the B200 path (embed -> align -> pos-enc -> tailored
    encoder -> fusion -> CTC loss + greedy), CUDA-graph replay included, against the oracle."""
    import copy

    from oracle import cases
    from tailored_avsr_b200.audiovisual_fusion.adaptive_audiovisual_fusion import AdaptiveAudioVisualFusion
    from tailored_avsr_b200.ctc.ctc import CTC
    from tailored_avsr_b200.encoder.audiovisual.tailored.encoder import TailoredEncoder
    from tailored_avsr_b200.pipeline import AVEncoderCTCPipeline
    ae, ve, sd = _build()
    cfg = dict(copy.deepcopy(cases.BASE_TAILORED), num_blocks=2, acoustic_use_attn=[True, False],
               visual_use_attn=[False, True])
    enc = TailoredEncoder(embed_pos_enc_layer_type="rel_pos", embed_rel_pos_type="latest", **cfg).eval()
    fusion = AdaptiveAudioVisualFusion(**cases.FUSION_DEFAULTS).eval()
    ctc = CTC(odim=37, encoder_output_size=256, dropout_rate=0.0).eval()
    sd.update(synth.fill_module(enc, seed=7))
    sd.update(synth.fill_module(fusion, seed=7, prefix="fusion."))
    sd.update(synth.fill_module(ctc, seed=7, prefix="ctc."))
    xa, la, xv, lv = gen_golden_embed.inputs()
    ys = synth.rand_targets(3, 10, 37, 3)
    yl = torch.tensor([10, 7, 4])
    # oracle composition
    o = _oracle(sd)
    with torch.no_grad():
        wa, wv = ref_path.tailored_encoder(o["audio_in"], o["pos"], o["audio_mask_aligned"], o["video_in"],
                                           o["pos"], o["video_mask_aligned"], sd, cfg)
        fused, olens, _ = ref_path.adaptive_av_fusion(wa, o["audio_mask_aligned"], wv,
                                                      o["video_mask_aligned"], sd, "fusion.")
        want_loss = float(ref_path.ctc_loss(fused, olens, ys, yl, sd, "ctc.ctc_lo"))
    pipe = AVEncoderCTCPipeline(enc.cuda(), fusion.cuda(), ctc.cuda(), acoustic_embed=ae.cuda(),
                                visual_embed=ve.cuda())
    dev_in = [t.cuda() for t in (xa, xv, la, lv, ys, yl)]
    pipe.run_device(*dev_in)                 # captures the graph
    res = pipe.run_device(*dev_in)           # replay
    torch.cuda.synchronize()
    got = res["encoder_out"].cpu()
    assert torch.equal(res["olens"].cpu().long(), olens.long())
    err = 0.0
    for b in range(3):
        n = int(olens[b])
        err = max(err, float((got[b, :n] - fused[b, :n]).abs().max() / fused.abs().max()))
    assert err <= 1e-3, err
    assert abs(float(res["loss"]) - want_loss) <= 2e-3 * abs(want_loss)   # loss on OUR encoder output
