"""Full-size checks at the bench workloads (BASELINE.json configs / SURVEY.md §8d C2, C4): parity
against the CPU oracle at the C2 size, and the size-independent properties of the path -
every operation is per-utterance, so an utterance's result may not depend on which
other utterances share its batch, on their order, or on the content of THEIR padding; the CTC loss
is the mean of the per-utterance losses."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(workload, batch=0, T=0):
    import bench
    bench.select_workload(types.SimpleNamespace(workload=workload, batch=batch, T=T))
    enc, fusion, ctc, sd = bench.build_modules()
    host, frames = bench.make_batch(0)
    return bench, enc.to(DEV).eval(), (fusion.to(DEV).eval() if fusion is not None else None), \
        ctc.to(DEV).eval(), sd, host


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def test_c2_full_size_parity_vs_oracle():
    """C2 (32 x 250 x 512, 12 two-branch blocks): encoder <= 1e-3, CTC loss <= 1e-4 given identical
    hs, greedy tokens exact given identical hs."""
    from oracle import ref_path
    bench, enc, _, ctc, sd, host = _setup("C2")
    feats, lens, ys, ylens = host
    with torch.no_grad():
        want, olens, _ = ref_path.branchformer_encoder(feats, lens, sd, bench.enc_cfg())
        got, golens, _ = enc(feats.to(DEV), lens.to(DEV))
        loss = ctc(want.to(DEV), olens.to(DEV), ys.to(DEV), ylens.to(DEV))
        toks = ctc.greedy_lists(want.to(DEV))
    assert torch.equal(golens.cpu().long(), olens.long())
    mx, fro = _rel(got, want)
    print(f"C2 full size: max-rel {mx:.3e} fro {fro:.3e}")
    assert mx <= 1e-3 and fro <= 1e-3, (mx, fro)
    ref_loss = float(ref_path.ctc_loss(want, olens, ys, ylens, sd, "ctc.ctc_lo"))
    assert abs(float(loss) - ref_loss) <= 1e-4 * abs(ref_loss)
    assert toks == ref_path.ctc_greedy(want, sd, "ctc.ctc_lo")


def test_c2_utterances_are_independent_of_batch_composition_and_order():
    bench, enc, _, ctc, sd, host = _setup("C2")
    feats, lens, ys, ylens = [t.to(DEV) for t in host]
    with torch.no_grad():
        full, _, _ = enc(feats, lens)
        full = full.clone()
        sub, _, _ = enc(feats[5:13].contiguous(), lens[5:13].contiguous())
        perm = torch.randperm(feats.shape[0], generator=torch.Generator().manual_seed(1)).to(DEV)
        shuf, _, _ = enc(feats[perm].contiguous(), lens[perm].contiguous())
        nll_all = ctc(full, lens, ys, ylens)
        ctc.reduce = False
        vec = ctc(full, lens, ys, ylens)
        ctc.reduce = True
        one = ctc(full[3:4].contiguous(), lens[3:4], ys[3:4], ylens[3:4])
    # the same rows go through the same arithmetic whatever tile they land in
    assert _rel(sub, full[5:13])[0] <= 1e-6
    assert _rel(shuf, full[perm])[0] <= 1e-6
    # loss = sum_b nll_b / B (ctc.py:62-66): the vector form sums to it; a batch of one gives nll_b
    assert abs(float(vec.sum()) - float(nll_all)) <= 1e-5 * abs(float(nll_all))
    assert abs(float(one) - float(vec[3]) * feats.shape[0]) <= 1e-5 * abs(float(one))


def test_c4_ragged_full_size_padding_of_one_utterance_does_not_leak_into_others():
    """C4 (tailored AV, 32 x 500 ragged): rewriting the padded tail of ONE utterance changes that
    utterance only (padded frames are computed densely and leak through the k=31 conv of their own
    utterance) - every other utterance must stay bit-identical."""
    from tailored_avsr_b200.pipeline import AVEncoderCTCPipeline
    bench, enc, fusion, ctc, sd, host = _setup("C4")
    pipe = AVEncoderCTCPipeline(enc, fusion, ctc, use_cuda_graph=False)
    a, v, la, lv, ys, ylens = [t.to(DEV) for t in host]
    b = int(torch.argmin(la))  # the shortest utterance has the longest padded tail
    assert int(la[b]) < a.shape[1]
    res0 = pipe.run_device(a, v, la, lv, ys, ylens)
    out0, loss0 = res0["encoder_out"].clone(), float(res0["loss"])
    a2 = a.clone()
    a2[b, int(la[b]):] = 7.0
    res1 = pipe.run_device(a2, v, la, lv, ys, ylens)
    out1 = res1["encoder_out"]
    others = [i for i in range(a.shape[0]) if i != b]
    assert torch.equal(out0[others], out1[others])
    assert not torch.equal(out0[b], out1[b])
    assert torch.isfinite(out1).all() and loss0 == loss0
    # valid-frame count and token lists come back for every utterance
    assert int(res1["olens"].sum()) == int(la.sum())
    assert res1["tokens"].shape[0] == a.shape[0]


def _av_oracle(bench, sd, host):
    """The CPU oracle on an AV workload of bench.py: encoder streams -> fusion -> (fused, olens)."""
    from oracle import cases, ref_path
    a, v, la, lv, ys, ylens = host
    cfg = bench.enc_cfg()
    T = a.shape[1]
    pos = ref_path.rel_pos_emb(T, 256)
    ma, mv = ref_path.make_valid_mask(la, T), ref_path.make_valid_mask(lv, T)
    if bench.WORKLOAD["kind"] == "tailored":
        ya, yv = ref_path.tailored_encoder(a, pos, ma, v, pos, mv, sd, cfg)[:2]
    else:
        ya, yv = ref_path.conventional_encoder(a, pos, ma, v, pos, mv, sd, cfg, cfg)[:2]
    fk = cases.FUSION_DEFAULTS
    fused, olens, _ = ref_path.adaptive_av_fusion(ya, ma, yv, mv, sd, "fusion.",
                                                  merge_method=fk["merge_method"], act=fk["activation_type"])
    return fused, olens


def _valid_rel(got, want, lens):
    g = torch.cat([got[b, : int(lens[b])] for b in range(got.shape[0])]).double().cpu()
    w = torch.cat([want[b, : int(lens[b])] for b in range(want.shape[0])]).double().cpu()
    return float((g - w).abs().max() / w.abs().max()), float((g - w).norm() / w.norm())


@pytest.mark.parametrize("workload,mode,tol", [("C3", "tf32", 1e-3), ("C3", "bf16", 5e-3),
                                              ("C4", "tf32", 1e-3), ("C4", "bf16", 5e-3),
                                              ("C2", "bf16", 5e-3)])
def test_full_size_parity_of_the_av_workloads_and_the_bf16_mode(workload, mode, tol):
    """Full-size parity vs the CPU oracle at the bench shapes the 2-3 block parity cases do not
    reach: C3 (AVSR conventional, 2 x 12 two-branch blocks + fusion, BASELINE.json configs[2]: the
    bf16 config) and C4 (tailored AV, 32 x 500 ragged, heterogeneous branches) in both compute
    modes, and C2 in bf16; valid frames only; the CTC loss of the device pipeline follows."""
    from oracle import ref_path
    from tailored_avsr_b200 import engine
    from tailored_avsr_b200.pipeline import AVEncoderCTCPipeline, EncoderCTCPipeline
    bench, enc, fusion, ctc, sd, host = _setup(workload)
    with torch.no_grad():
        if workload == "C2":
            feats, lens, ys, ylens = host
            want, olens, _ = ref_path.branchformer_encoder(feats, lens, sd, bench.enc_cfg())
            pipe = EncoderCTCPipeline(enc, ctc, use_cuda_graph=False)
        else:
            want, olens = _av_oracle(bench, sd, host)
            ys, ylens = host[4], host[5]
            pipe = AVEncoderCTCPipeline(enc, fusion, ctc, use_cuda_graph=False)
        with engine.use_compute_dtype(mode):
            res = pipe.run_device(*[t.to(DEV) for t in host])
        got = res["encoder_out"]
    assert torch.equal(res["olens"].cpu().long(), olens.long())
    mx, fro = _valid_rel(got, want, olens)
    print(f"FULLSIZE {workload} {mode}: max-rel {mx:.3e} fro {fro:.3e} (tol {tol:.0e})")
    assert mx <= tol and fro <= tol, (workload, mode, mx, fro)
    ref_loss = float(ref_path.ctc_loss(want, olens, ys, ylens, sd, "ctc.ctc_lo"))
    # the loss sees the encoder's own output here, so it carries the encoder tolerance
    assert abs(float(res["loss"]) - ref_loss) <= 20 * tol * abs(ref_loss), (float(res["loss"]), ref_loss)
