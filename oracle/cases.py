"""TEST INFRASTRUCTURE: the parity cases shared by oracle/gen_golden.py (which runs the REAL
reference on them), tests/test_oracle_cpu.py (oracle vs golden) and tests/test_parity_gpu.py
(CUDA path vs oracle and vs golden).  Every input and weight is regenerated from seeds
(oracle/synth.py), only the reference OUTPUTS are stored under tests/golden/."""
from __future__ import annotations

import copy

import torch

from . import synth

BASE_ENC = dict(
    output_size=256, attention_heads=4, linear_units=2048, num_blocks=12, cgmlp_linear_units=2048,
    cgmlp_conv_kernel=31, dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1,
    attn_branch_drop_rate=0.0, input_layer="conv2d", rel_pos_type="latest",
    pos_enc_layer_type="rel_pos", attention_layer_type="rel_selfattn",
    positionwise_layer_type="linear", ffn_activation_type="swish", merge_method="learned_ave",
    use_attn=True, use_cgmlp=True, macaron=True)

BASE_TAILORED = dict(
    output_size=256, attention_heads=4, linear_units=2048, num_blocks=12, dropout_rate=0.1,
    positional_dropout_rate=0.1, attention_dropout_rate=0.1, acoustic_branch_drop_rate=0.0,
    attention_layer_type="rel_selfattn", positionwise_layer_type="linear",
    ffn_activation_type="swish", cgmlp_linear_units=2048, cgmlp_conv_kernel=31,
    acoustic_use_attn=[False, True, True, True, False, True, False, True, False, True, True, True],
    visual_use_attn=[True, True, True, True, False, True, True, True, True, True, True, True],
    macaron=True, interctc_use_conditioning=False, audiovisual_interctc_conditioning=False)


def _enc(**kw):
    c = copy.deepcopy(BASE_ENC)
    c.update(kw)
    return c


# kind: "single" (MyBranchformerEncoder), "tailored", "conventional"
CASES = {
    # ASR-style conv2d front, learned_ave, ragged lengths (configs/ASR/branchformer_...english.yaml)
    "asr_small": dict(kind="single", input_size=80, cfg=_enc(num_blocks=3), B=3, Tin=203,
                      lens=[203, 150, 77], vocab=41, Lmax=12, seed=11),
    # VSR-style linear front (configs/VSR/conv3dresnet18_branchformer_...english.yaml, post-frontend)
    "vsr_small": dict(kind="single", input_size=512, cfg=_enc(num_blocks=2, input_layer="linear"),
                      B=2, Tin=70, lens=[70, 41], vocab=41, Lmax=20, seed=12),
    # tailored ASR: fixed_ave with pruned branches (configs/ASR/..._tailored.yaml:58-59)
    "asr_tailored_small": dict(kind="single", input_size=80,
                               cfg=_enc(num_blocks=3, merge_method="fixed_ave",
                                        cgmlp_weight=[1.0, 0.0, 0.0]),
                               B=2, Tin=163, lens=[163, 99], vocab=37, Lmax=10, seed=13),
    # the same branch pruning behind the linear front end (the training path's front ends are
    # linear / None): cgMLP-only, attention-only and a two-branch fixed_ave block in one stack
    "vsr_tailored_small": dict(kind="single", input_size=512,
                               cfg=_enc(num_blocks=3, input_layer="linear", merge_method="fixed_ave",
                                        cgmlp_weight=[1.0, 0.0, 0.4]),
                               B=2, Tin=61, lens=[61, 38], vocab=41, Lmax=9, seed=35),
    # dormant merges: concat, two-branch fixed_ave
    "concat_small": dict(kind="single", input_size=512,
                         cfg=_enc(num_blocks=2, input_layer="linear", merge_method="concat"),
                         B=2, Tin=50, lens=[50, 33], vocab=41, Lmax=8, seed=14),
    "fixed_ave_small": dict(kind="single", input_size=512,
                            cfg=_enc(num_blocks=2, input_layer="linear", merge_method="fixed_ave",
                                     cgmlp_weight=0.3),
                            B=2, Tin=50, lens=[50, 20], vocab=41, Lmax=8, seed=15),
    # unified tailored AV encoder (configs/AVSR/tailored_transformer+ctc_spanish.yaml:79-80 pattern)
    "av_tailored_small": dict(kind="tailored",
                              cfg=dict(BASE_TAILORED, num_blocks=3,
                                       acoustic_use_attn=[False, True, True],
                                       visual_use_attn=[True, False, True]),
                              B=2, T=60, lens=[60, 37], vocab=37, Lmax=15, seed=16),
    # conventional AV encoder: two independent stacks (configs/AVSR/conventional_...spanish.yaml)
    "av_conventional_small": dict(kind="conventional", cfg=_enc(num_blocks=2, input_layer=None),
                                  B=2, T=48, lens=[48, 31], vocab=37, Lmax=12, seed=17),
    # AdaptiveAudioVisualFusion behind the AV encoders: learned_ave fusion of
    # the two streams with DIFFERENT audio / video masks, CTC on the fused output
    # (avsr_espnet_model.py:467,678)
    "av_fusion_tailored": dict(kind="tailored",
                               cfg=dict(BASE_TAILORED, num_blocks=3,
                                        acoustic_use_attn=[True, False, True],
                                        visual_use_attn=[False, True, True]),
                               fusion=dict(merge_method="learned_ave"),
                               B=3, T=72, lens=[72, 50, 33], lens_video=[72, 41, 33], vocab=37, Lmax=14,
                               seed=21),
    "av_fusion_conventional": dict(kind="conventional", cfg=_enc(num_blocks=2, input_layer=None),
                                   fusion=dict(merge_method="learned_ave"),
                                   B=2, T=48, lens=[48, 31], lens_video=[48, 25], vocab=37, Lmax=12,
                                   seed=22),
    "av_fusion_fixed": dict(kind="tailored",
                            cfg=dict(BASE_TAILORED, num_blocks=2, acoustic_use_attn=[False, True],
                                     visual_use_attn=[True, True]),
                            fusion=dict(merge_method="fixed_ave", acoustic_weight=0.3),
                            B=2, T=40, lens=[40, 22], vocab=37, Lmax=9, seed=23),
    # audio-visual InterCTC: fused taps after blocks 1 and 2 + conditioning of both streams on the
    # fused posteriors (tailored/encoder.py:270-318)
    "av_tailored_interctc": dict(kind="tailored",
                                 cfg=dict(BASE_TAILORED, num_blocks=3,
                                          acoustic_use_attn=[False, True, True],
                                          visual_use_attn=[True, False, True],
                                          interctc_layer_idx=[1, 2], interctc_use_conditioning=True,
                                          audiovisual_interctc_conditioning=True),
                                 fusion=dict(merge_method="learned_ave"),
                                 B=2, T=56, lens=[56, 35], lens_video=[56, 30], vocab=37, Lmax=11,
                                 seed=24),
    "av_tailored_interctc_sep": dict(kind="tailored",
                                     cfg=dict(BASE_TAILORED, num_blocks=2,
                                              acoustic_use_attn=[True, True],
                                              visual_use_attn=[False, True],
                                              interctc_layer_idx=[1], interctc_use_conditioning=True,
                                              audiovisual_interctc_conditioning=False),
                                     fusion=dict(merge_method="learned_ave"),
                                     B=2, T=44, lens=[44, 28], vocab=37, Lmax=9, seed=25),
    # conventional AV encoder, layer-zipped InterCTC path with per-stream conditioning
    # (conventional/encoder.py:152-199)
    "av_conventional_interctc": dict(kind="conventional", cfg=_enc(num_blocks=2, input_layer=None),
                                     wrap=dict(interctc_layer_idx=[1], interctc_use_conditioning=True,
                                               audiovisual_interctc_conditioning=False),
                                     fusion=dict(merge_method="learned_ave"),
                                     B=2, T=40, lens=[40, 26], lens_video=[40, 21], vocab=37, Lmax=9,
                                     seed=26),
    # dormant InterCTC path: taps after blocks 1 and 2 + self-conditioning (encoder.py:378-401);
    # the model assigns conditioning_layer = Linear(V, d) (espnet_model.py:106-112)
    "asr_interctc_cond": dict(kind="single", input_size=512,
                              cfg=_enc(num_blocks=3, input_layer="linear", interctc_layer_idx=[1, 2],
                                       interctc_use_conditioning=True),
                              B=2, Tin=64, lens=[64, 39], vocab=41, Lmax=10, seed=19),
    # dormant early-exit path: forward(..., max_layer=1) runs blocks 0 and 1 of three, then
    # after_norm (encoder.py:370-374)
    "vsr_max_layer": dict(kind="single", input_size=512, cfg=_enc(num_blocks=3, input_layer="linear"),
                          B=2, Tin=58, lens=[58, 37], vocab=41, Lmax=9, seed=34, max_layer=1),
    # full-depth C1: 12 layers, B=8 x 10 s
    "asr_c1": dict(kind="single", input_size=80, cfg=_enc(), B=8, Tin=1001, lens=[1001] * 8,
                   vocab=41, Lmax=100, seed=1, stride_t=8, stride_d=4),
    # TF32 error-growth stress: same C1 shape at reduced batch, "hot" N(0,1/fan_in) weights (1.7x
    # larger branch outputs than default init); parity tolerance stated separately (3e-3)
    "asr_c1_hot": dict(kind="single", input_size=80, cfg=_enc(), B=2, Tin=1001, lens=[1001, 801],
                       vocab=41, Lmax=100, seed=2, stride_t=8, stride_d=4, hot=True, enc_tol=3e-3),
    # the longest sequence of the C5 sweep (T = 1500: 12 key tiles per attention row, 24 conv
    # segments) on a ragged pair, and the shortest ones (T = 5 < conv half-width 15; B = 1)
    "vsr_long1500": dict(kind="single", input_size=512, cfg=_enc(num_blocks=2, input_layer="linear"),
                         B=2, Tin=1500, lens=[1500, 777], vocab=41, Lmax=100, seed=31,
                         stride_t=4, stride_d=4),
    "vsr_tiny": dict(kind="single", input_size=512, cfg=_enc(num_blocks=2, input_layer="linear"),
                     B=1, Tin=5, lens=[5], vocab=41, Lmax=3, seed=32),
    "asr_shortest": dict(kind="single", input_size=80, cfg=_enc(num_blocks=2), B=2, Tin=23,
                         lens=[23, 9], vocab=41, Lmax=2, seed=33),
    # full-depth ragged VSR-like (C2/C4 flavour): padded variable-length batch, garbage in the pad
    "vsr_ragged12": dict(kind="single", input_size=512, cfg=_enc(input_layer="linear"), B=4, Tin=120,
                         lens=[120, 95, 64, 48], vocab=41, Lmax=30, seed=18, stride_t=4, stride_d=4),
}


def make_inputs(name: str):
    """Seeded inputs of a case: dict of tensors (CPU fp32)."""
    c = CASES[name]
    s = c["seed"]
    out = {}
    lens = torch.tensor(c["lens"], dtype=torch.int64)
    out["lens"] = lens
    if c["kind"] == "single":
        out["x"] = synth.randn((c["B"], c["Tin"], c["input_size"]), s)
    else:
        # block inputs post-embed, pre-pos-enc scaling is already applied by the model's
        # apply_pos_enc (avsr_espnet_model.py:447-448): x * sqrt(d); AV alignment pads with -1
        d = c["cfg"]["output_size"]
        a = synth.randn((c["B"], c["T"], d), s) * 4.0
        v = synth.randn((c["B"], c["T"], d), s + 1000) * 4.0
        for b, l in enumerate(c["lens"]):
            v[b, l:] = -1.0 * (d ** 0.5)
        out["audio"], out["video"] = a, v
        out["lens_video"] = torch.tensor(c.get("lens_video", c["lens"]), dtype=torch.int64)
    out["ys_pad"] = synth.rand_targets(c["B"], c["Lmax"], c["vocab"], s + 1)
    return out


FUSION_DEFAULTS = dict(input_size=256, output_size=256, hidden_units=2048,
                       audiovisual_layer_type="upsampling_positionwise", merge_method="learned_ave",
                       activation_type="swish", acoustic_weight=0.5, dropout_rate=0.1,
                       acoustic_branch_drop_rate=0.0)


def fusion_kwargs(name: str):
    """Constructor kwargs of the case's AdaptiveAudioVisualFusion, or None."""
    c = CASES[name]
    if "fusion" not in c:
        return None
    return dict(FUSION_DEFAULTS, **c["fusion"])


def target_lens(name: str, olens: torch.Tensor) -> torch.Tensor:
    """Target lengths: ~olens/3 clipped to Lmax; the LAST utterance gets Lmax (often infeasible for
    short inputs -> exercises zero_infinity)."""
    c = CASES[name]
    tl = (olens.long() // 3).clamp(1, c["Lmax"])
    tl[-1] = c["Lmax"]
    return tl
