"""bench.py's workload builders on the CPU: every SURVEY.md §8d workload constructs its drop-in
modules, the synthetic batch has the documented shapes, and the algorithmic FLOP counts are the
ones of SURVEY.md §8d (the figure `model_tflops` is computed from)."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _select(w, batch=0, T=0):
    return bench.select_workload(types.SimpleNamespace(workload=w, batch=batch, T=T))


def test_flops_per_frame_match_survey():
    _select("C2")
    assert bench.flops_per_frame() == pytest.approx(82.44e6, rel=1e-3)      # 12 two-branch blocks, T=250
    _select("C2", T=1500)
    assert bench.flops_per_frame() == pytest.approx(105.48e6, rel=1e-3)
    _select("C3")
    assert bench.flops_per_frame() == pytest.approx(2 * 82.44e6 + 4 * 256 * 2048, rel=1e-3)
    _select("C4")
    cfg = bench.enc_cfg()
    n_attn = sum(cfg["acoustic_use_attn"]) + sum(cfg["visual_use_attn"])
    mac = n_attn * (2359296 + 768 * 500) + (24 - n_attn) * 2915328 + 2 * 256 * 2048
    assert bench.flops_per_frame() == pytest.approx(2.0 * mac)


@pytest.mark.parametrize("name,n_in", [("C1", 4), ("C2", 4), ("C3", 6), ("C4", 6)])
def test_workload_batches_have_documented_shapes(name, n_in):
    w = _select(name, batch=3, T=40)
    host, frames = bench.make_batch(0)
    assert len(host) == n_in and w["B"] == 3 and w["T"] == 40
    if name == "C1":
        assert host[0].shape == (3, 4 * 40 + 5, 80) and frames == 3 * 40
    elif name == "C2":
        assert host[0].shape == (3, 40, 512) and frames == 3 * 40
    else:
        assert host[0].shape == host[1].shape == (3, 40, 256)
        assert frames == int(host[2].sum()) and int(host[2].max()) == 40
        if name == "C4":    # ragged: padded video tail carries the alignment value, targets ~ len / 3
            b = int(torch.argmin(host[2]))
            assert float(host[1][b, int(host[2][b]):].abs().min()) == 16.0 or int(host[2][b]) == 40
            assert torch.equal(host[5], (host[2] // 3).clamp(1, w["Lmax"]))
    host2, _ = bench.make_batch(1)
    assert not torch.equal(host[0], host2[0])       # every rank gets its own utterances


def test_workload_modules_construct_with_reference_parameter_counts():
    n = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    _select("C2")
    enc, fusion, ctc, sd = bench.build_modules()
    assert fusion is None and n(ctc) == 41 * 257
    _select("C3")
    enc, fusion, ctc, sd = bench.build_modules()
    assert fusion is not None and n(ctc) == 37 * 257
    assert n(fusion) == 4 * 257 + 2 * 256 * 2048 + 2048 + 256 + 512   # 4 x Linear(256,1), FFN, norm_final
    assert any(k.startswith("fusion.") for k in sd) and any(k.startswith("ctc.") for k in sd)
    _select("C4")
    enc, fusion, ctc, sd = bench.build_modules()
    assert n(enc) == 35625728                        # published tailored AV encoder size
    _select("C2")
