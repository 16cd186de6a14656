"""TEST INFRASTRUCTURE — CPU oracle ("port") of the tailored-avsr hot path.

A from-scratch, functional restatement in plain torch fp32 (CPU) of the Branchformer encoder stack
and the CTC scorer of david-gimeno/tailored-avsr.  It exists only to check the CUDA path: nothing
under tailored_avsr_b200/ may import it (only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do).

Pinning status.  The layer / encoder composition below is pinned against the reference's OWN
files (src/encoder/**, src/ctc/ctc.py) run unmodified on top of oracle/espnet_shim (see
oracle/gen_golden.py and tests/test_oracle_cpu.py; golden vectors in tests/golden/).  The leaf
arithmetic that lives in the un-vendored espnet==202402 (requirements.txt:1) is restated from its
published behaviour and the reference ships no tests or golden vectors, so
for those leaves: PARITY UNPINNED beyond the known answers (published parameter counts, rel-shift
index identity, torch.nn.CTCLoss, brute-force CTC path enumeration).

Every function takes a flat `state_dict`-style mapping of tensors and a key prefix, so the same
weights can be loaded into the reference modules, this oracle and the CUDA drop-in.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------
# leaves (espnet 202402 semantics)
# --------------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, sd: SD, prefix: str, eps: float = 1e-12) -> torch.Tensor:
    """espnet LayerNorm(eps=1e-12) (Appendix A.1); reference use: encoder_layer.py:102-109."""
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def activation(x: torch.Tensor, kind: str) -> torch.Tensor:
    """espnet get_activation (Appendix A.2); configs use "swish"."""
    if kind == "swish":
        return x * torch.sigmoid(x)
    if kind == "relu":
        return torch.relu(x)
    if kind == "gelu":
        return F.gelu(x)
    if kind == "tanh":
        return torch.tanh(x)
    raise ValueError(kind)


def feed_forward(x: torch.Tensor, sd: SD, prefix: str, act: str) -> torch.Tensor:
    """espnet PositionwiseFeedForward: w_2(act(w_1 x)); called at encoder_layer.py:194,314."""
    h = F.linear(x, sd[prefix + ".w_1.weight"], sd[prefix + ".w_1.bias"])
    return F.linear(activation(h, act), sd[prefix + ".w_2.weight"], sd[prefix + ".w_2.bias"])


def rel_pos_emb(T: int, d: int) -> torch.Tensor:
    """espnet RelPositionalEncoding table slice (Appendix A.3): row k <-> relative position T-1-k."""
    rel = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(rel * div)
    pe[:, 1::2] = torch.cos(rel * div)
    return pe.unsqueeze(0)


def rel_pos_mha(x: torch.Tensor, pos_emb: torch.Tensor, mask: torch.Tensor, sd: SD, prefix: str,
                heads: int) -> torch.Tensor:
    """espnet RelPositionMultiHeadedAttention.forward (Appendix A.4/A.5), called at
    encoder_layer.py:208.  mask: (B,1,T) bool, True = valid key.  The rel-shift is written as the
    explicit gather bd[i,j] = raw[i, T-1-i+j] (the identity Appendix A.5 states)."""
    B, T, D = x.shape
    dk = D // heads
    q = F.linear(x, sd[prefix + ".linear_q.weight"], sd[prefix + ".linear_q.bias"])
    k = F.linear(x, sd[prefix + ".linear_k.weight"], sd[prefix + ".linear_k.bias"])
    v = F.linear(x, sd[prefix + ".linear_v.weight"], sd[prefix + ".linear_v.bias"])
    q = q.view(B, T, heads, dk)
    k = k.view(B, T, heads, dk).transpose(1, 2)
    v = v.view(B, T, heads, dk).transpose(1, 2)
    p = F.linear(pos_emb, sd[prefix + ".linear_pos.weight"]).view(1, 2 * T - 1, heads, dk)
    p = p.transpose(1, 2)  # (1,h,2T-1,dk)
    q_u = (q + sd[prefix + ".pos_bias_u"]).transpose(1, 2)
    q_v = (q + sd[prefix + ".pos_bias_v"]).transpose(1, 2)
    ac = q_u @ k.transpose(-2, -1)
    raw = q_v @ p.transpose(-2, -1)  # (B,h,T,2T-1)
    idx = (T - 1 - torch.arange(T).unsqueeze(1) + torch.arange(T).unsqueeze(0))  # (T,T)
    bd = raw.gather(-1, idx.expand(B, heads, T, T))
    scores = (ac + bd) / math.sqrt(dk)
    inv = mask.unsqueeze(1).eq(0)  # (B,1,1,T)
    scores = scores.masked_fill(inv, torch.finfo(scores.dtype).min)
    attn = torch.softmax(scores, dim=-1).masked_fill(inv, 0.0)
    ctx = (attn @ v).transpose(1, 2).reshape(B, T, D)
    return F.linear(ctx, sd[prefix + ".linear_out.weight"], sd[prefix + ".linear_out.bias"])


def cgmlp(x: torch.Tensor, sd: SD, prefix: str, kernel: int) -> torch.Tensor:
    """espnet ConvolutionalGatingMLP (+CSGU, identity gate activation, no linear after conv)
    (Appendix A.6), called at encoder_layer.py:220.  The padding mask is ignored, as in espnet."""
    h = F.gelu(F.linear(x, sd[prefix + ".channel_proj1.0.weight"], sd[prefix + ".channel_proj1.0.bias"]))
    r, g = h.chunk(2, dim=-1)
    g = layer_norm(g, sd, prefix + ".csgu.norm")
    g = F.conv1d(g.transpose(1, 2), sd[prefix + ".csgu.conv.weight"], sd[prefix + ".csgu.conv.bias"],
                 padding=(kernel - 1) // 2, groups=g.shape[-1]).transpose(1, 2)
    return F.linear(r * g, sd[prefix + ".channel_proj2.weight"], sd[prefix + ".channel_proj2.bias"])


# --------------------------------------------------------------------------------------------------
# MyBranchformerEncoderLayer / MyBranchformerEncoder
# --------------------------------------------------------------------------------------------------
def _pool_weight(xb: torch.Tensor, mask: torch.Tensor, sd: SD, pool: str, wproj: str) -> torch.Tensor:
    """One branch of the learned_ave pooling, encoder_layer.py:242-258."""
    size = xb.shape[-1]
    score = F.linear(xb, sd[pool + ".weight"], sd[pool + ".bias"]).transpose(1, 2) / size ** 0.5
    min_value = float(np.finfo(np.float32).min)
    score = score.masked_fill(mask.eq(0), min_value)
    score = torch.softmax(score, dim=-1).masked_fill(mask.eq(0), 0.0)
    pooled = torch.matmul(score, xb).squeeze(1)
    return F.linear(pooled, sd[wproj + ".weight"], sd[wproj + ".bias"])


def branchformer_layer(x: torch.Tensor, pos_emb: torch.Tensor, mask: torch.Tensor, sd: SD,
                       prefix: str, *, heads: int = 4, kernel: int = 31, act: str = "swish",
                       merge_method: str = "learned_ave", cgmlp_weight: float = 0.5,
                       use_attn: bool = True, use_cgmlp: bool = True
                       ) -> Tuple[torch.Tensor, Optional[Tuple[torch.Tensor, torch.Tensor]]]:
    """Eval-mode restatement of MyBranchformerEncoderLayer.forward (encoder_layer.py:153-321).
    Returns the layer output and (weight_global, weight_local) when learned_ave is used."""
    has_attn = use_attn and not (merge_method == "fixed_ave" and use_cgmlp and cgmlp_weight == 1.0)
    has_mlp = use_cgmlp and not (merge_method == "fixed_ave" and use_attn and cgmlp_weight == 0.0)
    # macaron FFN (:191-194)
    x = x + 0.5 * feed_forward(layer_norm(x, sd, prefix + ".norm_ff_macaron"), sd,
                               prefix + ".feed_forward_macaron", act)
    x1 = x2 = x
    if has_attn:  # (:201-212)
        x1 = rel_pos_mha(layer_norm(x, sd, prefix + ".norm_mha"), pos_emb, mask, sd,
                         prefix + ".attn", heads)
    if has_mlp:  # (:215-224)
        x2 = cgmlp(layer_norm(x, sd, prefix + ".norm_mlp"), sd, prefix + ".cgmlp", kernel)
    weights = None
    mp_w, mp_b = sd.get(prefix + ".merge_proj.weight"), sd.get(prefix + ".merge_proj.bias")
    proj = (lambda t: F.linear(t, mp_w, mp_b)) if mp_w is not None else (lambda t: t)
    if has_attn and has_mlp:
        if merge_method == "concat":  # (:228-231)
            x = x + proj(torch.cat([x1, x2], dim=-1))
        elif merge_method == "learned_ave":  # (:232-293)
            om1 = _pool_weight(x1, mask, sd, prefix + ".pooling_proj1", prefix + ".weight_proj1")
            om2 = _pool_weight(x2, mask, sd, prefix + ".pooling_proj2", prefix + ".weight_proj2")
            mw = torch.softmax(torch.cat([om1, om2], dim=-1), dim=-1).unsqueeze(-1).unsqueeze(-1)
            w1, w2 = mw[:, 0], mw[:, 1]
            weights = (w1, w2)
            x = x + proj(w1 * x1 + w2 * x2)
        elif merge_method == "fixed_ave":  # (:294-299)
            x = x + proj((1.0 - cgmlp_weight) * x1 + cgmlp_weight * x2)
        else:
            raise RuntimeError(merge_method)
    elif has_mlp:  # (:303-304)
        x = x + proj(x2)
    else:  # (:305-306)
        x = x + proj(x1)
    # FFN + final norm (:311-316)
    x = x + 0.5 * feed_forward(layer_norm(x, sd, prefix + ".norm_ff"), sd, prefix + ".feed_forward", act)
    return layer_norm(x, sd, prefix + ".norm_final"), weights


def make_valid_mask(lens: torch.Tensor, T: Optional[int] = None) -> torch.Tensor:
    """~make_pad_mask(ilens)[:, None, :] (encoder.py:345): (B,1,T) bool, True = valid."""
    T = int(lens.max()) if T is None else T
    return (torch.arange(T).unsqueeze(0) < lens.long().unsqueeze(1)).unsqueeze(1)


def conv2d_subsample(x: torch.Tensor, mask: torch.Tensor, sd: SD, prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """espnet Conv2dSubsampling without the pos-enc step (Appendix A.8), encoder.py:149-155,364."""
    h = F.relu(F.conv2d(x.unsqueeze(1), sd[prefix + ".conv.0.weight"], sd[prefix + ".conv.0.bias"], stride=2))
    h = F.relu(F.conv2d(h, sd[prefix + ".conv.2.weight"], sd[prefix + ".conv.2.bias"], stride=2))
    b, c, t, f = h.shape
    h = F.linear(h.transpose(1, 2).contiguous().view(b, t, c * f),
                 sd[prefix + ".out.0.weight"], sd[prefix + ".out.0.bias"])
    return h, mask[:, :, :-2:2][:, :, :-2:2]


def branchformer_encoder(xs: torch.Tensor, ilens: torch.Tensor, sd: SD, cfg: dict, prefix: str = "",
                         max_layer: Optional[int] = None
                         ) -> Tuple[torch.Tensor, torch.Tensor, List[Optional[Tuple]]]:
    """Eval-mode restatement of MyBranchformerEncoder.forward (encoder.py:324-412), plain loop
    (:376) or the InterCTC loop (:378-401) when cfg["interctc_layer_idx"] is set, input_layer in
    {conv2d, linear, None}; `max_layer` is the early exit of :370-374 (only without InterCTC taps:
    blocks 0..max_layer run, then after_norm).  Returns (out, olens, per-layer merge weights); the tapped outputs are
    left in cfg-independent form on the function attribute `last_taps` [(idx, tensor)]."""
    d = cfg.get("output_size", 256)
    n = cfg.get("num_blocks", 12)
    masks = make_valid_mask(ilens, xs.shape[1])
    il = cfg.get("input_layer", "conv2d")
    if il == "conv2d":
        xs, masks = conv2d_subsample(xs, masks, sd, prefix + "embed")
    elif il == "linear":
        xs = F.linear(xs, sd[prefix + "embed.0.weight"], sd[prefix + "embed.0.bias"])
        xs = F.layer_norm(xs, (d,), sd[prefix + "embed.1.weight"], sd[prefix + "embed.1.bias"], 1e-5)
    elif il is not None:
        raise ValueError(f"oracle does not restate input_layer={il}")
    if il is not None:
        xs = xs * math.sqrt(d)  # RelPositionalEncoding (Appendix A.3)
        pos = rel_pos_emb(xs.shape[1], d)
    else:
        xs, pos = xs  # caller passes (x, pos_emb) like the AV wrappers do
    cw = cfg.get("cgmlp_weight", 0.5)
    cw = [cw] * n if isinstance(cw, float) else list(cw)
    weights = []
    taps = tuple(cfg.get("interctc_layer_idx", ()))
    tap_outs = []
    for l in range(n):
        xs, w = branchformer_layer(
            xs, pos, masks, sd, f"{prefix}encoders.{l}", heads=cfg.get("attention_heads", 4),
            kernel=cfg.get("cgmlp_conv_kernel", 31), act=cfg.get("ffn_activation_type", "relu"),
            merge_method=cfg.get("merge_method", "learned_ave"), cgmlp_weight=cw[l],
            use_attn=cfg.get("use_attn", True), use_cgmlp=cfg.get("use_cgmlp", True))
        weights.append(w)
        if not taps and max_layer is not None and 0 <= max_layer < n and l >= max_layer:
            break  # encoder.py:373-374
        if (l + 1) in taps:
            tap = layer_norm(xs, sd, prefix + "after_norm")  # :386-388
            tap_outs.append((l + 1, tap))
            if cfg.get("interctc_use_conditioning", False):  # :392-401
                post = torch.softmax(F.linear(tap, sd["ctc.ctc_lo.weight"], sd["ctc.ctc_lo.bias"]), 2)
                xs = xs + F.linear(post, sd[prefix + "conditioning_layer.weight"],
                                   sd[prefix + "conditioning_layer.bias"])
    branchformer_encoder.last_taps = tap_outs
    xs = layer_norm(xs, sd, prefix + "after_norm")
    return xs, masks.squeeze(1).sum(1), weights


# --------------------------------------------------------------------------------------------------
# audio-visual encoders
# --------------------------------------------------------------------------------------------------
def tailored_layer(audio, video, pos_a, pos_v, mask_a, mask_v, sd: SD, prefix: str, *,
                   a_attn: bool, v_attn: bool, heads: int = 4, kernel: int = 31, act: str = "swish"):
    """Eval-mode restatement of TailoredEncoderLayer.forward (tailored/encoder_layer.py:118-274):
    FFN-macaron / FFN / norm_final are shared; each stream owns one attention OR cgMLP branch."""
    outs = []
    for x, pos, mask, tag, use_attn in ((audio, pos_a, mask_a, "acoustic", a_attn),
                                        (video, pos_v, mask_v, "visual", v_attn)):
        x = x + 0.5 * feed_forward(layer_norm(x, sd, prefix + ".norm_ff_macaron"), sd,
                                   prefix + ".feed_forward_macaron", act)
        if use_attn:
            x = x + rel_pos_mha(layer_norm(x, sd, f"{prefix}.{tag}_norm_mha"), pos, mask, sd,
                                f"{prefix}.{tag}_attn", heads)
        else:
            x = x + cgmlp(layer_norm(x, sd, f"{prefix}.{tag}_norm_cgmlp"), sd,
                          f"{prefix}.{tag}_cgmlp", kernel)
        x = x + 0.5 * feed_forward(layer_norm(x, sd, prefix + ".norm_ff"), sd, prefix + ".feed_forward", act)
        outs.append(layer_norm(x, sd, prefix + ".norm_final"))
    return outs[0], outs[1]


def adaptive_av_fusion(audio, mask_a, video, mask_v, sd: SD, prefix: str = "fusion.", *,
                       merge_method: str = "learned_ave", acoustic_weight: float = 0.5,
                       act: str = "swish"):
    """Eval-mode restatement of AdaptiveAudioVisualFusion.forward
    (src/audiovisual_fusion/adaptive_audiovisual_fusion.py:113-211): per-modality masked softmax
    pooling over time -> 2-way softmax -> weighted average -> position-wise FFN (no residual, no
    pre-norm) -> norm_final.  Returns (fused, olens, acoustic_weight (B,1,1) or float)."""
    if merge_method == "learned_ave":
        wa = _pool_weight(audio, mask_a, sd, prefix + "acoustic_pooling_proj", prefix + "acoustic_weight_proj")
        wv = _pool_weight(video, mask_v, sd, prefix + "visual_pooling_proj", prefix + "visual_weight_proj")
        mw = torch.softmax(torch.cat([wa, wv], dim=-1), dim=-1).unsqueeze(-1).unsqueeze(-1)
        w_a, w_v = mw[:, 0], mw[:, 1]
    elif merge_method == "fixed_ave":
        w_a, w_v = acoustic_weight, 1.0 - acoustic_weight
    else:
        raise NotImplementedError(merge_method)
    z = feed_forward(w_a * audio + w_v * video, sd, prefix + "audiovisual_layer", act)
    fused = layer_norm(z, sd, prefix + "norm_final")
    olens = torch.logical_or(mask_a, mask_v).squeeze(1).sum(1)
    return fused, olens, w_a


def tailored_encoder(audio, pos_a, mask_a, video, pos_v, mask_v, sd: SD, cfg: dict, prefix: str = "",
                     fusion=None, ctc_softmax=None):
    """Eval-mode restatement of TailoredEncoder.forward (tailored/encoder.py:221-332).  With
    `interctc_layer_idx` in cfg, `fusion(a, mask_a, v, mask_v) -> fused` produces the tapped
    audio-visual outputs (:270-289) and `ctc_softmax` the conditioning posteriors (:291-318).
    Returns (audio, video) or (audio, video, taps)."""
    audio = audio + sd[prefix + "modality_encoding.weight"][0]
    video = video + sd[prefix + "modality_encoding.weight"][1]
    idx = list(cfg.get("interctc_layer_idx", []) or [])
    taps = []
    for l in range(cfg.get("num_blocks", 12)):
        audio, video = tailored_layer(
            audio, video, pos_a, pos_v, mask_a, mask_v, sd, f"{prefix}encoders.{l}",
            a_attn=cfg["acoustic_use_attn"][l], v_attn=cfg["visual_use_attn"][l],
            heads=cfg.get("attention_heads", 4), kernel=cfg.get("cgmlp_conv_kernel", 31),
            act=cfg.get("ffn_activation_type", "swish"))
        if l + 1 in idx:
            ea = layer_norm(audio, sd, prefix + "after_norm")
            ev = layer_norm(video, sd, prefix + "after_norm")
            eav = fusion(ea, mask_a, ev, mask_v)
            taps.append((l + 1, eav))
            if cfg.get("interctc_use_conditioning", False):
                if cfg.get("audiovisual_interctc_conditioning", False):
                    ca = cv = ctc_softmax(eav)
                else:
                    ca, cv = ctc_softmax(ea), ctc_softmax(ev)
                cw, cb = sd[prefix + "conditioning_layer.weight"], sd[prefix + "conditioning_layer.bias"]
                audio = audio + F.linear(ca, cw, cb)
                video = video + F.linear(cv, cw, cb)
    a_out = layer_norm(audio, sd, prefix + "after_norm")
    v_out = layer_norm(video, sd, prefix + "after_norm")
    if idx:
        return a_out, v_out, taps
    return a_out, v_out


def conventional_encoder(audio, pos_a, mask_a, video, pos_v, mask_v, sd: SD, cfg_a: dict, cfg_v: dict,
                         prefix: str = ""):
    """Eval-mode restatement of ConventionalEncoder.forward (conventional/encoder.py:116-217):
    two independent Branchformer stacks (embed is None), each with its own after_norm."""
    ca = dict(cfg_a, input_layer=None)
    cv = dict(cfg_v, input_layer=None)
    B, T = audio.shape[:2]
    la = mask_a.squeeze(1).sum(1)
    lv = mask_v.squeeze(1).sum(1)
    a, _, wa = _encoder_with_mask((audio, pos_a), mask_a, sd, ca, prefix + "acoustic_encoder.")
    v, _, wv = _encoder_with_mask((video, pos_v), mask_v, sd, cv, prefix + "visual_encoder.")
    return a, v, wa, wv


def conventional_encoder_interctc(audio, pos_a, mask_a, video, pos_v, mask_v, sd: SD, cfg: dict,
                                  wrap: dict, fusion, ctc_softmax, prefix: str = ""):
    """Layer-zipped ConventionalEncoder.forward with audio-visual InterCTC taps
    (conventional/encoder.py:152-199).  `wrap` holds the wrapper's interctc_layer_idx /
    interctc_use_conditioning / audiovisual_interctc_conditioning.  Returns (audio, video, taps)."""
    n = cfg.get("num_blocks", 12)
    cw = cfg.get("cgmlp_weight", 0.5)
    cw = [cw] * n if isinstance(cw, float) else list(cw)
    xs = [audio, video]
    taps = []
    for l in range(n):
        for k, (pos, mask, tag) in enumerate(((pos_a, mask_a, "acoustic_encoder."),
                                              (pos_v, mask_v, "visual_encoder."))):
            xs[k], _ = branchformer_layer(
                xs[k], pos, mask, sd, f"{prefix}{tag}encoders.{l}", heads=cfg.get("attention_heads", 4),
                kernel=cfg.get("cgmlp_conv_kernel", 31), act=cfg.get("ffn_activation_type", "relu"),
                merge_method=cfg.get("merge_method", "learned_ave"), cgmlp_weight=cw[l],
                use_attn=cfg.get("use_attn", True), use_cgmlp=cfg.get("use_cgmlp", True))
        if l + 1 in wrap["interctc_layer_idx"]:
            ea = layer_norm(xs[0], sd, prefix + "acoustic_encoder.after_norm")
            ev = layer_norm(xs[1], sd, prefix + "visual_encoder.after_norm")
            eav = fusion(ea, mask_a, ev, mask_v)
            taps.append((l + 1, eav))
            if wrap.get("interctc_use_conditioning", False):
                if wrap.get("audiovisual_interctc_conditioning", False):
                    ca_ = cv_ = ctc_softmax(eav)
                else:
                    ca_, cv_ = ctc_softmax(ea), ctc_softmax(ev)
                w_, b_ = sd[prefix + "conditioning_layer.weight"], sd[prefix + "conditioning_layer.bias"]
                xs[0] = xs[0] + F.linear(ca_, w_, b_)
                xs[1] = xs[1] + F.linear(cv_, w_, b_)
    return (layer_norm(xs[0], sd, prefix + "acoustic_encoder.after_norm"),
            layer_norm(xs[1], sd, prefix + "visual_encoder.after_norm"), taps)


def _encoder_with_mask(x_pos, masks, sd, cfg, prefix):
    xs, pos = x_pos
    n = cfg.get("num_blocks", 12)
    cw = cfg.get("cgmlp_weight", 0.5)
    cw = [cw] * n if isinstance(cw, float) else list(cw)
    weights = []
    for l in range(n):
        xs, w = branchformer_layer(
            xs, pos, masks, sd, f"{prefix}encoders.{l}", heads=cfg.get("attention_heads", 4),
            kernel=cfg.get("cgmlp_conv_kernel", 31), act=cfg.get("ffn_activation_type", "relu"),
            merge_method=cfg.get("merge_method", "learned_ave"), cgmlp_weight=cw[l],
            use_attn=cfg.get("use_attn", True), use_cgmlp=cfg.get("use_cgmlp", True))
        weights.append(w)
    return layer_norm(xs, sd, prefix + "after_norm"), masks.squeeze(1).sum(1), weights


# --------------------------------------------------------------------------------------------------
# AVSR embedding layers (src/embedding_for_avsr/default.py) + the model's temporal alignment
# --------------------------------------------------------------------------------------------------
def avsr_embed_layer(xs: torch.Tensor, ilens: torch.Tensor, sd: SD, prefix: str, input_layer: str):
    """DefaultEmbeddingLayerForAVSR.apply_embed_layer (default.py:139-153), eval mode.
    conv2d = espnet Conv2dSubsamplingWOPosEnc (two 3x3 stride-2 convs + ReLU, Linear, no pos-enc);
    linear = Linear + torch LayerNorm (eps 1e-5).  Returns (x (B,T,d), masks (B,1,T))."""
    masks = make_valid_mask(ilens, xs.shape[1])
    if input_layer == "conv2d":
        h = F.relu(F.conv2d(xs.unsqueeze(1), sd[prefix + "embed.conv.0.weight"],
                            sd[prefix + "embed.conv.0.bias"], stride=2))
        h = F.relu(F.conv2d(h, sd[prefix + "embed.conv.2.weight"], sd[prefix + "embed.conv.2.bias"],
                            stride=2))
        b, c, t, f = h.shape
        x = F.linear(h.transpose(1, 2).contiguous().view(b, t, c * f), sd[prefix + "embed.out.weight"],
                     sd[prefix + "embed.out.bias"])
        return x, masks[:, :, :-2:2][:, :, :-2:2]
    if input_layer == "linear":
        x = F.linear(xs, sd[prefix + "embed.0.weight"], sd[prefix + "embed.0.bias"])
        x = F.layer_norm(x, (x.shape[-1],), sd[prefix + "embed.1.weight"], sd[prefix + "embed.1.bias"], 1e-5)
        return x, masks
    raise NotImplementedError(input_layer)


def audiovisual_alignment(a, ma, v, mv, ignore_id: float = -1.0):
    """ESPnetAVSRModel.audiovisual_alignment (src/models/avsr_espnet_model.py:512-541): the shorter
    stream is padded with ignore_id (masks with False) up to the longer one."""
    pad = a.shape[1] - v.shape[1]
    if pad < 0:
        a = F.pad(a, (0, 0, 0, -pad, 0, 0), value=ignore_id)
        ma = F.pad(ma, (0, -pad), value=False)
    elif pad > 0:
        v = F.pad(v, (0, 0, 0, pad, 0, 0), value=ignore_id)
        mv = F.pad(mv, (0, pad), value=False)
    return a, ma, v, mv


def rel_pos_enc(x: torch.Tensor):
    """espnet RelPositionalEncoding.forward in eval: (x * sqrt(d), pos_emb (1, 2T-1, d))."""
    return x * math.sqrt(x.shape[-1]), rel_pos_emb(x.shape[1], x.shape[-1])


# --------------------------------------------------------------------------------------------------
# CTC (src/ctc/ctc.py)
# --------------------------------------------------------------------------------------------------
def interctc_residual(x: torch.Tensor, sd: SD, prefix: str = "") -> Tuple[torch.Tensor, torch.Tensor]:
    """InterCTCResidualModule.forward (src/ctc/interctc_residual_module.py:11-16)."""
    logits = F.linear(x, sd[prefix + "proj_1.weight"], sd[prefix + "proj_1.bias"])
    y = x + F.linear(torch.softmax(logits, -1), sd[prefix + "proj_2.weight"], sd[prefix + "proj_2.bias"])
    return y, logits


def ctc_log_softmax(hs: torch.Tensor, sd: SD, prefix: str = "ctc_lo") -> torch.Tensor:
    """CTC.log_softmax (ctc.py:170-178)."""
    return F.log_softmax(F.linear(hs, sd[prefix + ".weight"], sd[prefix + ".bias"]), dim=2)


def ctc_nll_numpy(logp: np.ndarray, target: Sequence[int], blank: int = 0) -> float:
    """Log-domain CTC alpha recursion for ONE utterance in float64 (what torch.nn.CTCLoss computes,
    Appendix A.10; reference call site ctc.py:60-61).  logp: (T,V) log-probabilities."""
    T = logp.shape[0]
    ext = [blank]
    for c in target:
        ext += [int(c), blank]
    S = len(ext)
    if T == 0:
        return 0.0 if len(target) == 0 else float("inf")
    alpha = np.full(S, -np.inf)
    alpha[0] = logp[0, ext[0]]
    if S > 1:
        alpha[1] = logp[0, ext[1]]
    for t in range(1, T):
        new = np.full(S, -np.inf)
        for s in range(S):
            cands = [alpha[s]]
            if s >= 1:
                cands.append(alpha[s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                cands.append(alpha[s - 2])
            m = max(cands)
            if m > -np.inf:
                new[s] = m + math.log(sum(math.exp(c - m) for c in cands)) + logp[t, ext[s]]
        alpha = new
    tail = [alpha[S - 1]] + ([alpha[S - 2]] if S > 1 else [])
    m = max(tail)
    if m == -np.inf:
        return float("inf")
    return -(m + math.log(sum(math.exp(c - m) for c in tail)))


def ctc_loss(hs: torch.Tensor, hlens: torch.Tensor, ys_pad: torch.Tensor, ys_lens: torch.Tensor,
             sd: SD, prefix: str = "ctc_lo", reduce: bool = True, zero_infinity: bool = True
             ) -> torch.Tensor:
    """CTC.forward with dropout_rate = 0 and ctc_type="builtin" (ctc.py:133-158, 58-69):
    loss = sum_b nll_b / B (or the vector nll_b / B when reduce is False)."""
    logp = ctc_log_softmax(hs, sd, prefix)
    B = hs.shape[0]
    ys_true = torch.cat([ys_pad[i, : int(l)] for i, l in enumerate(ys_lens)])
    nll = F.ctc_loss(logp.transpose(0, 1), ys_true, hlens.long(), ys_lens.long(), blank=0,
                     reduction="none", zero_infinity=zero_infinity)
    return nll.sum() / B if reduce else nll / B


def ctc_greedy(hs: torch.Tensor, sd: SD, prefix: str = "ctc_lo", lens: Optional[torch.Tensor] = None,
               blank: int = 0) -> List[List[int]]:
    """ctc.argmax (ctc.py:180-188) + groupby collapse + blank removal (espnet_model.py:590-592,
    maskctc_model.py:287-291).  lens=None collapses over all Tmax frames like _calc_ctc_loss."""
    ids = torch.argmax(F.linear(hs, sd[prefix + ".weight"], sd[prefix + ".bias"]), dim=2)
    out = []
    for b in range(ids.shape[0]):
        seq = ids[b, : (int(lens[b]) if lens is not None else ids.shape[1])].tolist()
        out.append([t for i, t in enumerate(seq) if t != blank and (i == 0 or seq[i - 1] != t)])
    return out


LOGZERO = -1e10


def ctc_prefix_init(logp: np.ndarray, blank: int = 0) -> np.ndarray:
    """Initial forward variables of the empty prefix (Appendix A.9): r^n = logzero,
    r^b_t = cumsum_t logp[t, blank].  Returns (T,2)."""
    T = logp.shape[0]
    r = np.full((T, 2), LOGZERO, dtype=np.float64)
    r[:, 1] = np.cumsum(logp[:, blank])
    return r


def _lae(a: float, b: float) -> float:
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def ctc_prefix_score(logp: np.ndarray, r_prev: np.ndarray, prefix: Sequence[int], blank: int, eos: int
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """One CTCPrefixScoreTH step for one hypothesis (Appendix A.9; call site
    src/inference/asr_inference.py:142).  Returns (r_new (V,T,2), log_psi (V,)) — log_psi is the
    absolute prefix log-probability; the scorer's score is log_psi - log_psi(prefix)."""
    T, V = logp.shape
    plen = len(prefix)
    last = prefix[-1] if plen > 0 else -1
    start = max(plen, 1)
    r_sum = np.array([_lae(r_prev[t, 0], r_prev[t, 1]) for t in range(T)])
    r_new = np.full((V, T, 2), LOGZERO, dtype=np.float64)
    log_psi = np.full(V, LOGZERO, dtype=np.float64)
    for c in range(V):
        if c == blank:
            continue
        phi = r_prev[:, 1] if c == last else r_sum
        rn, rb = LOGZERO, LOGZERO
        if plen == 0:
            rn = logp[0, c]
        r_new[c, start - 1] = (rn, rb)
        psi = rn
        for t in range(start, T):
            nn = _lae(rn, phi[t - 1]) + logp[t, c]
            nb = _lae(rn, rb) + logp[t, blank]
            psi = _lae(psi, phi[t - 1] + logp[t, c])
            rn, rb = nn, nb
            r_new[c, t] = (rn, rb)
        log_psi[c] = psi
    log_psi[eos] = r_sum[T - 1]
    return r_new, log_psi


def ctc_beam_search(logp: np.ndarray, beam: int, eos: int, blank: int = 0, ctc_weight: float = 1.0,
                    length_bonus: float = 0.0, maxlen: Optional[int] = None, nbest: int = 1):
    """CTC-only beam search with the semantics of espnet's BatchBeamSearch driven by the `ctc`
    partial scorer (the object built at src/inference/asr_inference.py:142,276-303): per output
    position every running hypothesis is extended by every token with the prefix-score difference,
    the `beam` best (hypothesis, token) pairs survive, hypotheses ending in <eos> move to the ended
    list, at the last position only <eos> may follow.  Plain Python over ctc_prefix_score; returns
    [(tokens, score)] best first (the checker of tailored_avsr_b200.ctc.beam_search)."""
    T, V = logp.shape
    maxlen = T if maxlen is None else maxlen
    running = [([], ctc_prefix_init(logp, blank), 0.0, 0.0)]       # (prefix, r, log_psi, score)
    ended = []
    for i in range(maxlen):
        cands = []
        for prefix, r, psi_prev, score in running:
            r_new, log_psi = ctc_prefix_score(logp, r, prefix, blank, eos)
            for c in range(V):
                if c == blank or (i == maxlen - 1 and c != eos):
                    continue
                sc = LOGZERO if log_psi[c] <= LOGZERO / 2 else log_psi[c] - psi_prev
                cands.append((score + ctc_weight * sc + length_bonus, prefix, c, r_new[c], log_psi[c]))
        cands.sort(key=lambda t: -t[0])
        running = []
        for total, prefix, c, r_c, psi_c in cands[:beam]:
            if total <= -1e9:
                continue
            if c == eos:
                ended.append((prefix, total))
            else:
                running.append((prefix + [c], r_c, psi_c, total))
        if not running:
            break
        if length_bonus <= 0 and len(ended) >= nbest:
            best_end = sorted((s_ for _, s_ in ended), reverse=True)[nbest - 1]
            if best_end >= max(t[3] for t in running):
                break
    ended.sort(key=lambda t: -t[1])
    return ended[:nbest]
