#!/usr/bin/env python
"""bench.py — Branchformer encoder + CTC valid frames/s on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload = "C2"): BASELINE.json configs[1] / SURVEY.md §8d C2 — VSR video-only
Branchformer (input_layer=linear, 12 two-branch learned_ave blocks, d=256, 4 heads, FFN 2048, cgMLP
2048, k=31) + CTC (V=41): post-frontend features (32, 250, 512) per GPU, targets (32, 100).
One step = encoder forward + CTC loss + greedy CTC decode of one batch (a validation step of the
reference: src/models/espnet_model.py:397-402,578-593).  Synthetic inputs, seeded default-init-like
weights (oracle/synth.py).

value      = valid frames / s, inputs resident in HBM, CUDA-graph replay, CUDA-event timed per
             step with an L2 flush (256 MB memset) between steps, summed over K steps, max over ranks.
e2e        = same metric through EncoderCTCPipeline.run_stream() with pinned HOST inputs: H2D
             copies of features / lengths / targets and D2H of loss + greedy tokens of every step
             inside the timed region (copy of batch n+1 overlapped with the kernels of batch n);
             e2e.blocking_call_value = the same through one blocking EncoderCTCPipeline.run() per batch.
roofline   = dominant kernel group of the step (CUDA events around every op of one eager step).
cpu_baseline / --impl reference = the CPU oracle port (oracle/ref_path.py; the reference's own
             modules cannot travel to the GPU box: espnet is not installable) on the
             host cores with all threads, on the FULL batch of the same workload (median of the steps).
train      = (extra object of the same line) a training step of the same workload: encoder forward
             in grad mode (training.py: one autograd node per block) + CTC loss normalised by the
             GLOBAL batch + backward + bucketed gradient all-reduce (parallel.GradBucketReducer,
             launched from gradient hooks DURING backward) + no optimizer; the exposed (not
             overlapped) all-reduce time is reported.  `--mode train` makes it the headline value.
strong     = (extra object) strong scaling: a FIXED global batch of 64 utterances split over the N
             ranks (the inference step of the same workload).
gpu_eager_baseline = (N = 1, informational) the plain-PyTorch port of the reference path
             (oracle/ref_path.py) run eagerly on the same B200, fp32 and TF32-allowed, i.e. the
             cuBLAS / ATen-composed path BASELINE.md §3 names as the bar to beat.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "branchformer_encoder_ctc_valid_frames_per_sec"
UNIT = "frames/s"

# SURVEY.md §8d workloads.  B = utterances PER GPU, T = encoder frames.  C2 is the bench line
# (BASELINE.json configs[1]); the others are selected with --workload for the sweep / parity-size
# measurements (tools/sweep.py), never silently.
WORKLOADS = {
    "C1": dict(kind="single", front="conv2d", B=8, T=249, feat=80, vocab=41, Lmax=100,
               desc="ASR audio-only, conv2d front end on (B, 4T+5, 80) log-mel"),
    "C2": dict(kind="single", front="linear", B=32, T=250, feat=512, vocab=41, Lmax=100,
               desc="VSR video-only, linear front end on (B, T, 512) post-frontend features"),
    "C3": dict(kind="conventional", B=8, T=250, vocab=37, Lmax=100,
               desc="AVSR conventional: two 12-block stacks + AdaptiveAudioVisualFusion, "
                    "(B, T, 256) block inputs per stream"),
    "C4": dict(kind="tailored", B=32, T=500, vocab=37, Lmax=166, ragged=True,
               desc="AVSR tailored: heterogeneous per-layer branches "
                    "(configs/AVSR/tailored_transformer+ctc_spanish.yaml:79-80), padded "
                    "variable-length batch, AdaptiveAudioVisualFusion"),
}
WORKLOAD = dict(WORKLOADS["C2"], name="C2", layers=12)


def select_workload(args):
    global WORKLOAD
    w = dict(WORKLOADS[args.workload], name=args.workload, layers=12)
    if args.batch:
        w["B"] = args.batch
    if args.T:
        w["T"] = args.T
        if w["name"] == "C4":
            w["Lmax"] = max(1, args.T // 3)
    WORKLOAD = w
    return w


def enc_cfg():
    from oracle import cases
    w = WORKLOAD
    if w["kind"] == "single":
        return dict(cases.BASE_ENC, input_layer=w["front"])
    if w["kind"] == "conventional":
        return dict(cases.BASE_ENC, input_layer=None)
    return dict(cases.BASE_TAILORED)


def flops_per_frame():
    """Algorithmic FLOP per valid frame, 12 blocks; AV workloads count both
    streams plus the fusion FFN."""
    w = WORKLOAD
    T = w["T"]
    if w["kind"] == "single":
        return 2.0 * 12 * (3243008 + 768 * T)
    if w["kind"] == "conventional":
        return 2.0 * (2 * 12 * (3243008 + 768 * T) + 2 * 256 * 2048)
    cfg = enc_cfg()
    mac = 0
    for l in range(12):
        for use_attn in (cfg["acoustic_use_attn"][l], cfg["visual_use_attn"][l]):
            mac += (2359296 + 768 * T) if use_attn else 2915328
    return 2.0 * (mac + 2 * 256 * 2048)


def build_modules():
    """(encoder, fusion-or-None, ctc) drop-in modules on the CPU with the seeded synthetic weights,
    plus the flat state dict the CPU oracle reads."""
    from oracle import cases, synth
    from tailored_avsr_b200.audiovisual_fusion.adaptive_audiovisual_fusion import AdaptiveAudioVisualFusion
    from tailored_avsr_b200.ctc.ctc import CTC
    from tailored_avsr_b200.encoder.audiovisual.conventional.encoder import ConventionalEncoder
    from tailored_avsr_b200.encoder.audiovisual.tailored.encoder import TailoredEncoder
    from tailored_avsr_b200.encoder.branchformer.encoder import MyBranchformerEncoder
    w = WORKLOAD
    cfg = enc_cfg()
    fusion = None
    if w["kind"] == "single":
        enc = MyBranchformerEncoder(input_size=w["feat"], **cfg)
    elif w["kind"] == "conventional":
        sub = {k: v for k, v in cfg.items() if k != "output_size"}
        enc = ConventionalEncoder(input_size=256,
                                  acoustic_encoder_conf=dict(sub, encoder_class_type="branchformer"),
                                  visual_encoder_conf=dict(sub, encoder_class_type="branchformer"),
                                  output_size=256)
    else:
        enc = TailoredEncoder(embed_pos_enc_layer_type="rel_pos", embed_rel_pos_type="latest", **cfg)
    ctc = CTC(odim=w["vocab"], encoder_output_size=256, dropout_rate=0.0)
    sd = synth.fill_module(enc, seed=0)
    sd.update(synth.fill_module(ctc, seed=0, prefix="ctc."))
    if w["kind"] != "single":
        fusion = AdaptiveAudioVisualFusion(**cases.FUSION_DEFAULTS)
        sd.update(synth.fill_module(fusion, seed=0, prefix="fusion."))
    return enc, fusion, ctc, sd


def make_batch(rank: int):
    """Host tensors of one step, in the positional order of the pipeline's run()."""
    from oracle import synth
    w = WORKLOAD
    B, T = w["B"], w["T"]
    ys = synth.rand_targets(B, w["Lmax"], w["vocab"], 4 + 100 * rank)
    if w["kind"] == "single":
        Tin = T if w["front"] == "linear" else 4 * T + 5
        feats = synth.randn((B, Tin, w["feat"]), 3 + 100 * rank)
        lens = torch.full((B,), Tin, dtype=torch.int64)
        ylens = torch.full((B,), w["Lmax"], dtype=torch.int64)
        return [feats, lens, ys, ylens], B * T
    a = synth.randn((B, T, 256), 4 + 100 * rank) * 4.0
    v = synth.randn((B, T, 256), 5 + 100 * rank) * 4.0
    if w.get("ragged"):
        lens = synth.rand_lens(B, T, 6 + 100 * rank)
        for b in range(B):
            v[b, int(lens[b]):] = -16.0  # the AV alignment pad (-1 before x sqrt(d))
        ylens = (lens // 3).clamp(1, w["Lmax"])
    else:
        lens = torch.full((B,), T, dtype=torch.int64)
        ylens = torch.full((B,), w["Lmax"], dtype=torch.int64)
    return [a, v, lens, lens.clone(), ys, ylens], int(lens.sum())


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# --------------------------------------------------------------------------------------------------
def cpu_oracle_arm(steps: int, warmup: int, sample_B: int):
    """Times the CPU oracle port on `sample_B` utterances of the workload per step; returns
    (frames/s from the MEDIAN step, median ms per step, threads)."""
    from oracle import cases, ref_path
    w = WORKLOAD
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = enc_cfg()
    _, _, _, sd = build_modules()  # parameter containers only (CPU)
    batch, _ = make_batch(0)
    batch = [t[:sample_B] for t in batch]

    def step():
        with torch.no_grad():
            if w["kind"] == "single":
                feats, lens, ys, ylens = batch
                out, olens, _ = ref_path.branchformer_encoder(feats, lens, sd, cfg)
            else:
                a, v, la, lv, ys, ylens = batch
                T = a.shape[1]
                pos = ref_path.rel_pos_emb(T, 256)
                ma, mv = ref_path.make_valid_mask(la, T), ref_path.make_valid_mask(lv, T)
                if w["kind"] == "tailored":
                    ya, yv = ref_path.tailored_encoder(a, pos, ma, v, pos, mv, sd, cfg)
                else:
                    ya, yv, _, _ = ref_path.conventional_encoder(a, pos, ma, v, pos, mv, sd, cfg, cfg)
                fk = cases.FUSION_DEFAULTS
                out, olens, _ = ref_path.adaptive_av_fusion(ya, ma, yv, mv, sd, "fusion.",
                                                            merge_method=fk["merge_method"],
                                                            act=fk["activation_type"])
            loss = ref_path.ctc_loss(out, olens, ys, ylens, sd, "ctc.ctc_lo")
            toks = ref_path.ctc_greedy(out, sd, "ctc.ctc_lo")
        return float(loss), toks, int(olens.sum())

    for _ in range(warmup):
        step()
    times, frames = [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        frames = step()[2]
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return frames / med, med * 1e3, torch.get_num_threads()


CPU_STEPS, CPU_WARMUP = 5, 1   # the cpu_baseline leg of the GPU arm (full batch, median)


def run_reference_arm(args):
    """--impl reference: the CPU oracle port on the FULL batch of the workload (the GPU arm's
    config), args.steps steps after args.warmup warm-ups (capped so the run stays within minutes:
    a C2 step takes ~1 s on 16 cores), value from the median step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample_B = WORKLOAD["B"]
    steps = max(1, min(args.steps, 30))
    warm = max(1, min(args.warmup, 5))
    fps, ms, cores = cpu_oracle_arm(steps, warm, sample_B)
    sample = (f"the full {WORKLOAD['name']} batch per step ({sample_B}x{WORKLOAD['T']} frames), median of "
              f"{steps} steps after {warm} warm-up, fp32, torch {torch.__version__} with {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "batch_per_gpu": sample_B, "T": WORKLOAD["T"],
                   "feat": WORKLOAD.get("feat", 256), "layers": 12, "vocab": WORKLOAD["vocab"],
                   "target_len": WORKLOAD["Lmax"],
                   "note": "CPU oracle port of the reference path (espnet not installable: the "
                           "reference's own modules cannot run on the box); rank 0 only"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
def op_cost(name, shapes, extra, eb=4):
    """(algorithmic flops, algorithmic bytes) of one op call from its tensor shapes; eb = bytes per
    element of the operand-only tensors (4 in tf32 mode, 2 in bf16 mode), the residual stream and
    the encoder output are fp32 in every mode."""
    if name == "gemm_bias_act":
        (M, K), (N, _) = shapes[0], shapes[1]
        return 2.0 * M * N * K, eb * (M * K + N * K + M * N)
    if name == "gemm_rowln":
        (M, K), (N, _) = shapes[0], shapes[1]
        dual = 2 if "x2" in extra else 1
        outs = ("lnA" in extra) + ("lnB" in extra)
        return 2.0 * M * N * K * dual, eb * (dual * M * K + N * K + outs * M * N) + 4.0 * M * N * (
            1 + ("residual" in extra))
    if name == "ffn_fused":
        M = shapes[0][0]
        H, D = shapes[1]
        # algorithmic bytes: xn in, up to 2 LayerNorm outputs (operand storage); residual in and
        # the main output (fp32); weights once
        return 4.0 * M * H * D, eb * (M * D * 3 + 2 * H * D) + 4.0 * M * D * 2
    if name == "relpos_attn":
        M, C = shapes[0]
        T = (shapes[1][0] + 1) // 2
        d = C // 3
        return 2.0 * 3 * T * d * M, eb * (M * C + M * d)
    if name == "csgu":
        M, C = shapes[0]
        return 2.0 * M * (C // 2) * 31, eb * (M * C + M * C // 2)
    if name == "layernorm":
        M, D = shapes[0]
        return 8.0 * M * D, 8.0 * M * D
    if name == "ctc_head":
        M, D = shapes[0]
        V = shapes[1][0]
        return 2.0 * M * D * V, 4.0 * (M * D + M * V)
    return 0.0, 0.0


def load_ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full`
    capture) keyed by compute mode and op group, written by tools/ncu_traffic.py from a capture of
    THIS build (the file records the library digest it was taken on)."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if not os.path.exists(path):
        return {}, None
    with open(path) as f:
        d = json.load(f)
    return d, path


KERNEL_SOURCES = {   # the files that define each measured kernel group (tools/ncu_traffic.py writes
    # their digest at capture time; a late change elsewhere in the library does not stale them)
    "ffn_fused": ["ffn_sm100.cu", "ffn_sm100.cuh", "gemm_sm100.cuh", "ptx.cuh", "host.h"],
    "gemm": ["gemm_sm100.cu", "gemm_sm100.cuh", "ptx.cuh", "host.h"],
    "relpos_attn": ["attention_sm100.cu", "ptx.cuh", "host.h"],
    "csgu": ["rowops.cu", "ptx.cuh", "host.h"],
    "merge_scores": ["rowops.cu", "ptx.cuh", "host.h"],
    "ctc": ["ctc.cu", "ptx.cuh", "host.h"],
}


def kernel_source_digest(group):
    import hashlib
    files = KERNEL_SOURCES.get(group)
    if files is None:
        return None
    h = hashlib.sha256()
    try:
        for f in files:
            with open(os.path.join(ROOT, "tailored_avsr_b200", "csrc", f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    except OSError:
        return None
    return h.hexdigest()[:16]


def lib_digest():
    try:
        with open(os.path.join(ROOT, "tailored_avsr_b200", "libtavsr_sm100.so.digest")) as f:
            return f.read().strip()[:16]
    except OSError:
        return None


def measure_dense_peaks(dev):
    """cuBLAS dense peaks measured the MEASURED_PEAKS.json way (torch.matmul 8192^3, best of 10,
    CUDA events): bf16 and TF32 (fp32 storage with allow_tf32).  The TF32 figure is the roofline
    denominator of the tf32 mode; the driver's file only holds the bf16 one."""
    out = {}
    n = 8192
    for name, dt, tf32 in (("bf16", torch.bfloat16, False), ("tf32", torch.float32, True)):
        a = torch.randn(n, n, device=dev, dtype=dt)
        b = torch.randn(n, n, device=dev, dtype=dt)
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        try:
            best = 1e9
            for _ in range(3):
                torch.matmul(a, b)
            for _ in range(10):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name + "_tflops"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        del a, b
    return out


def profile_step(pipe, batch_dev, peaks, dtype, live_peaks):
    """One eager step with CUDA events around every op: per-op-group time shares + roofline of the
    dominant group."""
    from tailored_avsr_b200 import ops
    eb = 2 if dtype == "bf16" else 4
    recs = []
    torch.cuda.synchronize()
    ops.set_profiler(recs)
    try:
        with torch.no_grad():
            # park the GPU behind a ~40 ms spin so the whole step (~125 launches + their event pairs)
            # is enqueued before the first kernel starts: the event deltas are then kernel
            # durations, not host launch latency
            torch.cuda._sleep(80_000_000)
            pipe._step(*batch_dev)
        torch.cuda.synchronize()
    finally:
        ops.set_profiler(None)
    groups = {}
    total = 0.0
    for name, shapes, extra, e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        key = f"{name}{list(shapes[:2])}" + ("+dual" if "x2" in extra else "")
        fl, by = op_cost(name, shapes, extra, eb)
        g = groups.setdefault(key, {"ms": 0.0, "n": 0, "flops": fl, "bytes": by, "name": name})
        g["ms"] += ms
        g["n"] += 1
        total += ms
    top_key = max(groups, key=lambda k: groups[k]["ms"])
    top = groups[top_key]
    avg_s = top["ms"] / top["n"] * 1e-3
    tensor_bound = top["name"] in ("gemm_bias_act", "gemm_rowln", "relpos_attn", "ffn_fused")
    if tensor_bound:
        achieved = top["flops"] / avg_s / 1e12
        peak = peaks["bf16_tflops"]
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": None,
                "peak_note": "measured dense bf16 burst (MEASURED_PEAKS.json); in the tf32 mode the "
                             "kernel's own ceiling is the TF32 dense peak, measured live below"}
        if live_peaks:
            roof["peak_bf16_tflops_live"] = live_peaks.get("bf16_tflops")
            roof["peak_tf32_tflops_live"] = live_peaks.get("tf32_tflops")
            if dtype != "bf16" and live_peaks.get("tf32_tflops"):
                roof["frac_of_tf32_peak"] = achieved / live_peaks["tf32_tflops"]
    else:
        achieved = top["bytes"] / avg_s / 1e9
        peak = peaks["hbm_gbs"]
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None}
    # DRAM traffic per launch of that kernel from an ncu --set full capture of this build
    table, path = load_ncu_traffic()
    entry = table.get(dtype, {}).get(top_key) if table else None
    if entry is not None:
        roof["traffic"] = entry
        roof["traffic_source"] = os.path.relpath(path, ROOT)
        roof["traffic_build_digest"] = table.get("lib_digest")
        roof["traffic_build_matches"] = table.get("lib_digest") == lib_digest()
        # the capture also records a digest of the source files that define each kernel group: the
        # figure is current as long as THOSE files are byte-identical, whatever else changed
        grp = next((g for g in KERNEL_SOURCES if top_key.startswith(g)), None)
        rec = (table.get("kernel_source_digests") or {}).get(grp)
        roof["traffic_kernel_sources_match"] = (rec is not None and rec == kernel_source_digest(grp))
        roof["traffic_captured_on_commit"] = table.get("captured_on_commit")
    roof["algorithmic_bytes"] = top["bytes"]
    roof["algorithmic_flops"] = top["flops"]
    roof["kernel"] = top_key
    roof["avg_launch_us"] = avg_s * 1e6
    roof["share_of_step"] = top["ms"] / total if total > 0 else None
    shares = {k: round(v["ms"] / total, 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
    per_launch_us = {k: round(v["ms"] / v["n"] * 1e3, 2) for k, v in sorted(groups.items(), key=lambda kv: -kv[1]["ms"])}
    return roof, shares, total, per_launch_us


# --------------------------------------------------------------------------------------------------
# extra legs of the GPU arm
# --------------------------------------------------------------------------------------------------
def trainable_workload():
    return True   # every workload's modules train (InterCTC, unused by the workloads, excepted)


def train_forward(enc, fusion, ctc, batch):
    """Grad-mode forward of the trainable workloads: (nll vector of the local utterances).  C2:
    encoder + CTC; C3 / C4: ConventionalEncoder (two stacks) / TailoredEncoder +
    AdaptiveAudioVisualFusion + CTC on the fused stream (avsr_espnet_model.py:467,678)."""
    w = WORKLOAD
    if w["kind"] == "single":
        feats, lens, ys, ylens = batch
        out, olens, _ = enc(feats, lens)
        return ctc(out, olens, ys, ylens)
    from tailored_avsr_b200.espnet_compat import RelPositionalEncoding
    a, v, lens_a, lens_v, ys, ylens = batch
    B, T, d = a.shape
    pos = _TRAIN_POS.setdefault(d, RelPositionalEncoding(d, 0.0)).pos_emb(T, a.device)
    ar = torch.arange(T, device=a.device)[None, :]
    ma, mv = (ar < lens_a[:, None]).unsqueeze(1), (ar < lens_v[:, None]).unsqueeze(1)
    ya, _, yv, _, _ = enc((a, pos), ma, (v, pos), mv)
    y, olens = fusion(ya, ma, yv, mv)
    return ctc(y, olens, ys, ylens)


_TRAIN_POS = {}


def train_leg(args, enc, ctc, rank, world, dev, dist, steps, warmup, train_mode=True, fusion=None):
    """Training step of the workload (single-stream workloads): grad-mode encoder forward + CTC loss
    / global batch + backward + overlapped bucketed gradient all-reduce, no optimizer.  The modules
    are in train() mode: every dropout site of the reference is active at the configured rate (0.1
    like the shipped YAMLs; `train_mode=False` times the same step in eval() mode).  Inputs come
    from pinned host memory every step and the loss is read back (inside the timed region).
    Returns a dict (rank 0) or None."""
    from tailored_avsr_b200 import engine, ops, parallel
    w = WORKLOAD
    if not trainable_workload():
        return {"unavailable": "no training path for this workload"}
    prev = engine.compute_dtype()
    engine.set_compute_dtype("tf32")      # the training path stores fp32 and multiplies in TF32
    cpu_threads = torch.get_num_threads()
    if os.environ.get("TAVSR_BENCH_KEEP_THREADS", "0") != "1":
        # the eager step is bound by its Python-side launches; torch's intra-op CPU pool (idle here)
        # measurably slows that thread when left at the core count, torchrun sets it to 1 anyway
        torch.set_num_threads(1)
    enc.train(train_mode)
    if fusion is not None:
        fusion.train(train_mode)
    host, frames = make_batch(rank)
    host = [t.pin_memory() for t in host]
    params = list(enc.parameters()) + list(ctc.parameters())
    if fusion is not None:
        params += list(fusion.parameters())
    for p in params:
        p.requires_grad_(True)
    red = parallel.GradBucketReducer(params, bucket_mb=25.0, overlap=True)
    Bg = w["B"] * world
    ctc.reduce = False

    host_split = [0.0, 0.0, 0.0, 0.0]   # forward, loss, backward, finish (host enqueue seconds)

    def step():
        for p in params:
            p.grad = None
        t0 = time.perf_counter()
        batch = [t.to(dev, non_blocking=True) for t in host]
        vec = train_forward(enc, fusion, ctc, batch) * w["B"]   # nll_b of the local utterances
        t1 = time.perf_counter()
        loss = vec.sum() / Bg                               # ctc.py:62-66 with the GLOBAL batch
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        e_b = torch.cuda.Event(enable_timing=True)
        e_b.record()
        n = red.finish()
        e_c = torch.cuda.Event(enable_timing=True)
        e_c.record()
        t4 = time.perf_counter()
        for i, dt in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            host_split[i] += dt
        return loss, n, e_b, e_c

    for _ in range(max(1, warmup)):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    l0 = ops.launch_count()
    evs, exposed, in_bwd = [], [], []
    host_enqueue_s = 0.0
    host_split[:] = [0.0, 0.0, 0.0, 0.0]
    for _ in range(steps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        red.launched_in_backward = 0
        th = time.perf_counter()
        loss, n_coll, e_b, e_c = step()
        host_enqueue_s += time.perf_counter() - th
        loss_host = float(loss.detach())                    # D2H read of the step's result
        e1.record()
        evs.append((e0, e1))
        exposed.append((e_b, e_c))
        in_bwd.append(red.launched_in_backward)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    launches = ops.launch_count() - l0
    ms = sum(a.elapsed_time(b) for a, b in evs)
    exp_ms = sum(a.elapsed_time(b) for a, b in exposed)
    t = torch.tensor([ms, exp_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, exp_ms = float(t[0]), float(t[1])
    red.remove_hooks()
    ctc.reduce = True
    for p in params:
        p.grad = None
        p.requires_grad_(False)
    enc.eval()
    if fusion is not None:
        fusion.eval()
    engine.set_compute_dtype(prev)
    torch.set_num_threads(cpu_threads)
    if rank != 0:
        return None
    nbytes = sum(red.bucket_bytes())
    rates = sorted({float(m.p) for m in enc.modules() if isinstance(m, torch.nn.Dropout)} |
                   {float(getattr(m, "dropout_rate", 0.0)) for m in enc.modules()})
    return {
        "value": frames * steps * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
        "steps": steps, "warmup": max(1, warmup), "dtype": "tf32 (fp32 storage)",
        "module_mode": ("train(): dropout active at every site of the reference, rates "
                        f"{rates}; stochastic depth / branch drop as configured (0)") if train_mode
                       else "eval(): no dropout",
        "step": "grad-mode encoder forward + CTC loss / global batch + backward + bucketed gradient "
                "all-reduce (no optimizer); host inputs copied in and the loss read back every step",
        "global_batch": Bg, "scaling": "weak", "loss": loss_host,
        "kernel_launches_per_step": launches // steps,
        "host_enqueue_ms_per_step": host_enqueue_s / steps * 1e3,
        "host_enqueue_split_ms": {k: round(v / steps * 1e3, 2) for k, v in
                                  zip(("forward + loss kernels", "loss normalisation", "backward", "finish"),
                                      host_split)},
        "allreduce": {"bytes_per_step": nbytes if world > 1 else 0, "buckets": len(red.buckets),
                      "bucket_mb": 25.0, "collectives_per_step": n_coll,
                      "launched_during_backward": in_bwd[-1] if in_bwd else 0,
                      "exposed_ms_per_step": exp_ms / steps,
                      "note": "exposed = CUDA-event time from the end of backward() to the last "
                              "bucket's sum being back in .grad (waits + unpack copies), max over ranks"},
    }


def train_graph_leg(args, enc, ctc, rank, world, dev, dist, steps, warmup, fusion=None):
    """The same training step captured ONCE into a CUDA graph and replayed (the eager step is bound
    by ~1500 Python-side launches).  First choice: the bucketed all-reduce is captured too, launched
    by the gradient hooks during the captured backward, so the replayed graph overlaps the NCCL
    transfers with the backward kernels exactly like the eager step.  If NCCL capture is refused the
    graph holds forward + loss + backward and the all-reduce runs after the replay (not overlapped);
    the result says which.  Returns a dict (rank 0) or None."""
    from tailored_avsr_b200 import engine, parallel
    w = WORKLOAD
    if not trainable_workload():
        return None
    prev = engine.compute_dtype()
    engine.set_compute_dtype("tf32")
    enc.train()                  # dropout active, masks drawn inside the capture (graph-safe generator)
    out = None
    params = list(enc.parameters()) + list(ctc.parameters())
    if fusion is not None:
        fusion.train()
        params += list(fusion.parameters())
    host, frames = make_batch(rank)
    host = [t.pin_memory() for t in host]
    static = [t.to(dev) for t in host]
    Bg = w["B"] * world
    ctc.reduce = False

    def attempt(overlap):
        for p in params:
            p.requires_grad_(True)
            p.grad = None
        red = parallel.GradBucketReducer(params, bucket_mb=25.0, overlap=overlap)

        def fwd_bwd():
            red.launched_in_backward = 0
            vec = train_forward(enc, fusion, ctc, static) * w["B"]
            loss = vec.sum() / Bg
            loss.backward()
            n = red.finish() if overlap else 0
            return loss, n

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                for p in params:
                    p.grad = None
                fwd_bwd()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss, n_in_graph = fwd_bwd()
        in_bwd = red.launched_in_backward

        def step():
            for d_, h_ in zip(static, host):
                d_.copy_(h_, non_blocking=True)
            graph.replay()
            e_b = torch.cuda.Event(enable_timing=True)
            e_b.record()
            n = n_in_graph if overlap else red.reduce()
            e_c = torch.cuda.Event(enable_timing=True)
            e_c.record()
            return n, e_b, e_c

        for _ in range(max(1, warmup)):
            step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        evs, exposed = [], []
        for _ in range(steps):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            n_coll, e_b, e_c = step()
            loss_host = float(loss.detach())
            e1.record()
            evs.append((e0, e1))
            exposed.append((e_b, e_c))
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        exp_ms = sum(a.elapsed_time(b) for a, b in exposed)
        t = torch.tensor([ms, exp_ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, exp_ms = float(t[0]), float(t[1])
        res = {"value": frames * steps * world / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
               "steps": steps, "loss": loss_host, "global_batch": Bg,
               "step": ("ONE CUDA-graph replay of forward + loss + backward + the bucketed gradient "
                        "all-reduce (NCCL nodes captured where the gradient hooks launched them: "
                        "overlapped with the backward)") if overlap else
                       ("CUDA-graph replay of forward + loss + backward, then the bucketed gradient "
                        "all-reduce (not overlapped)"),
               "io": "host inputs copied in and the loss read back every step",
               "allreduce": {"collectives_per_step": n_coll, "captured_in_graph": bool(overlap),
                             "launched_during_backward": in_bwd if overlap else 0,
                             "exposed_ms_per_step": None if overlap else exp_ms / steps,
                             "bytes_per_step": sum(red.bucket_bytes()) if world > 1 else 0}}
        red.remove_hooks()
        del graph
        return res

    try:
        try:
            out = attempt(overlap=world > 1)
        except Exception as e:  # noqa: BLE001
            if world == 1:
                raise
            torch.cuda.synchronize()
            first = f"{type(e).__name__}: {e}"[:200]
            out = attempt(overlap=False)
            out["overlapped_capture_failed"] = first
    except Exception as e:  # noqa: BLE001
        out = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.synchronize()
    finally:
        ctc.reduce = True
        for p in params:
            p.grad = None
            p.requires_grad_(False)
        enc.eval()
        if fusion is not None:
            fusion.eval()
        engine.set_compute_dtype(prev)
    return out if rank == 0 else None


def strong_leg(args, enc, fusion, ctc, rank, world, dev, dist, steps, warmup, global_batch=64):
    """Strong scaling: a FIXED global batch split over the ranks (utterance sharding), inference
    step of the workload through CUDA-graph replay.  Returns a dict (rank 0) or None."""
    from tailored_avsr_b200.pipeline import AVEncoderCTCPipeline, EncoderCTCPipeline
    global WORKLOAD
    if global_batch % world != 0:
        return {"unavailable": f"global batch {global_batch} does not split over {world} ranks"}
    saved = WORKLOAD
    WORKLOAD = dict(saved, B=global_batch // world)
    try:
        host, frames = make_batch(rank)
        batch_dev = [t.to(dev) for t in host]
        pipe = (EncoderCTCPipeline(enc, ctc) if fusion is None
                else AVEncoderCTCPipeline(enc, fusion, ctc))
        for _ in range(max(3, warmup)):
            pipe.run_device(*batch_dev)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        evs = []
        for _ in range(steps):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe.replay_static()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms, float(frames)], device=dev, dtype=torch.float64)
        if dist is not None:
            tm = t[:1].clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            tf = t[1:].clone()
            dist.all_reduce(tf, op=dist.ReduceOp.SUM)
            ms, total = float(tm[0]), float(tf[0])
        else:
            total = float(frames)
        del flush
    finally:
        WORKLOAD = saved
    if rank != 0:
        return None
    return {"value": total * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
            "global_batch": global_batch, "batch_per_gpu": global_batch // world, "scaling": "strong",
            "steps": steps, "note": "fixed global batch split by utterance over the ranks; below ~32 "
                                    "utterances per GPU a step sits on the launch-latency floor"}


def gpu_eager_baseline(dev, steps=5):
    """The plain-PyTorch port of the reference path (oracle/ref_path.py: F.linear / matmul /
    F.conv1d / F.layer_norm, i.e. cuBLAS + ATen kernels) run eagerly on the same GPU: encoder
    forward + torch CTC loss + argmax.  Informational (BASELINE.md §3's comparison bar)."""
    import torch.nn.functional as F
    from oracle import ref_path
    w = WORKLOAD
    if w["kind"] != "single":
        return {"unavailable": "single-stream workloads only"}
    cfg = enc_cfg()
    _, _, _, sd = build_modules()
    sd = {k: v.to(dev) for k, v in sd.items()}
    feats, lens, ys, ylens = [t.to(dev) for t in make_batch(0)[0]]
    out = {}
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        for name, tf32 in (("fp32", False), ("tf32_allowed", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32

            def step():
                # torch.device(dev): the port's factory calls (arange, zeros, the positional table)
                # land on the GPU too
                with torch.no_grad(), torch.device(dev):
                    y, olens, _ = ref_path.branchformer_encoder(feats, lens, sd, cfg)
                    lp = F.log_softmax(F.linear(y, sd["ctc.ctc_lo.weight"], sd["ctc.ctc_lo.bias"]), -1)
                    loss = F.ctc_loss(lp.transpose(0, 1), ys, olens, ylens, reduction="sum",
                                      zero_infinity=True) / y.shape[0]
                    return loss, lp.argmax(-1), int(w["B"] * w["T"])

            for _ in range(2):
                step()
            torch.cuda.synchronize()
            times = []
            for _ in range(steps):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                _, _, frames = step()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            med = statistics.median(times)
            out[name] = {"value": frames / (med * 1e-3), "unit": UNIT, "ms_per_step": med}
    except Exception as e:  # noqa: BLE001
        out["error"] = f"{type(e).__name__}: {e}"[:200]
    # the same port with its launch overhead removed: the encoder forward (all but ~1 % of the work)
    # captured into a CUDA graph, TF32 allowed.  torch's CTC loss reads its length tensors on the host
    # and cannot be captured, so this figure is encoder-only and is given beside the eager encoder-only
    # time; a port that syncs inside the encoder reports why it cannot be captured.
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True

        def enc_only():
            with torch.no_grad(), torch.device(dev):
                return ref_path.branchformer_encoder(feats, lens, sd, cfg)[0]

        def timed(fn):
            ts = []
            for _ in range(steps):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return statistics.median(ts)

        frames = int(w["B"] * w["T"])
        enc_only()
        torch.cuda.synchronize()
        eager_ms = timed(enc_only)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            enc_only()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            enc_only()
        graph.replay()
        torch.cuda.synchronize()
        graph_ms = timed(graph.replay)
        out["tf32_allowed_encoder_only"] = {
            "eager_ms_per_step": eager_ms, "cuda_graph_ms_per_step": graph_ms,
            "cuda_graph_value": frames / (graph_ms * 1e-3), "unit": UNIT}
        del graph
    except Exception as e:  # noqa: BLE001
        out["tf32_allowed_encoder_only"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    out["what"] = ("oracle/ref_path.py (plain torch ops: cuBLAS GEMMs, ATen elementwise / softmax / "
                   "LayerNorm, cuDNN depthwise conv) eager on the same GPU, encoder + torch CTC loss + "
                   "argmax, median of %d steps" % steps)
    return out


def run_gpu_arm(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    from tailored_avsr_b200 import engine, ops
    from tailored_avsr_b200.pipeline import AVEncoderCTCPipeline, EncoderCTCPipeline
    engine.set_compute_dtype(args.dtype)

    w = WORKLOAD
    peaks, peak_src = load_peaks()
    enc, fusion, ctc, _ = build_modules()
    enc, ctc = enc.to(dev).eval(), ctc.to(dev).eval()
    if fusion is None:
        pipe = EncoderCTCPipeline(enc, ctc, use_cuda_graph=not args.no_graph)
    else:
        pipe = AVEncoderCTCPipeline(enc, fusion.to(dev).eval(), ctc, use_cuda_graph=not args.no_graph)

    host, frames_per_step = make_batch(rank)
    host = [t.pin_memory() for t in host]
    batch_dev = [t.to(dev) for t in host]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # ---- warm-up (also captures the CUDA graph) ----
    for _ in range(max(args.warmup, 3)):
        res = pipe.run_device(*batch_dev)
    for _ in pipe.run_stream(host for _ in range(max(args.warmup, 3))):  # staging buffers, copy stream
        pass
    pipe.run(*host)
    torch.cuda.synchronize()
    loss_ref = float(res["loss"])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    # ---- timed: device-resident ----
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if args.no_graph:
            pipe.run_device(*batch_dev)
        else:
            pipe.replay_static()
        e1.record()
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    eager_launches = ops.launch_count() - launches0

    # ---- timed: end to end through the public API with host inputs ----
    # (a) one blocking call per batch: pipe.run(host batch) -> host results
    barrier()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pipe.run(*host)
        e1.record()
        evs.append((e0, e1))
    barrier()
    e2e_call_ms = sum(a.elapsed_time(b) for a, b in evs)
    assert abs(float(out["loss"]) - loss_ref) <= 1e-5 * abs(loss_ref), "e2e loss differs"
    # (b) the streaming form of the same API: pipe.run_stream(batches) overlaps the H2D copy of
    # batch n+1 with the kernels of batch n.  One event pair around all K steps; every step's
    # inputs cross PCIe and every step's loss / tokens are read back inside it; the L2 flush
    # between steps is INSIDE the timed region here (conservative).
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for out in pipe.run_stream(host for _ in range(args.steps)):
        flush.zero_()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    assert abs(float(out["loss"]) - loss_ref) <= 1e-5 * abs(loss_ref), "e2e loss differs"

    # kernels per step: count them on one eager (non-graph) step
    l0 = ops.launch_count()
    with torch.no_grad():
        pipe._step(*batch_dev)
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - l0
    del eager_launches

    t = torch.tensor([dev_ms, e2e_ms, e2e_call_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_call_ms = float(t[0]), float(t[1]), float(t[2])
    total_frames = frames_per_step * args.steps * world

    # ---- the other compute mode on the same pipeline (device-resident graph replay, short) ----
    other_mode = None
    if not args.no_extra:
        other = "tf32" if args.dtype != "tf32" else "bf16"
        engine.set_compute_dtype(other)
        for _ in range(3):
            pipe.run_device(*batch_dev)
        barrier()
        key = tuple((tuple(t.shape), t.dtype) for t in batch_dev) + (other,)
        evs = []
        for _ in range(min(args.steps, 10)):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            pipe.replay_static(key)
            e1.record()
            evs.append((e0, e1))
        barrier()
        oms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([oms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        other_mode = {"dtype": other, "ms_per_step": float(t[0]) / len(evs),
                      "value": frames_per_step * len(evs) * world / (float(t[0]) * 1e-3), "unit": UNIT,
                      "note": "same workload and pipeline in the other compute mode, device-resident "
                              "graph replay (parity: tf32 <= 1e-3, bf16 <= 5e-3 vs the fp32 oracle)"}
        engine.set_compute_dtype(args.dtype)

    # ---- extra legs: strong scaling and the training step (every rank takes part) ----
    strong = train = None
    if not args.no_extra:
        strong = strong_leg(args, enc, fusion, ctc, rank, world, dev, dist, steps=min(args.steps, 10),
                            warmup=3)
        train = train_leg(args, enc, ctc, rank, world, dev, dist, steps=args.train_steps, warmup=2,
                          fusion=fusion)
        train_graph = train_graph_leg(args, enc, ctc, rank, world, dev, dist, steps=args.train_steps,
                                      warmup=2, fusion=fusion)
        train_eval = train_leg(args, enc, ctc, rank, world, dev, dist, steps=max(2, args.train_steps // 2),
                               warmup=1, train_mode=False, fusion=fusion)
        if train is not None and train_eval is not None and "ms_per_step" in train_eval:
            train["eval_mode_ms_per_step"] = train_eval["ms_per_step"]
        if train is not None and train_graph is not None:
            if "value" in train_graph and "value" in train:
                # the graph replay is the training step to quote (the eager step is bound by the
                # Python-side enqueue of ~1500 launches); the eager, hook-driven step stays beside it
                eager = train
                train = dict(train_graph)
                for k in ("dtype", "module_mode", "warmup", "scaling", "kernel_launches_per_step"):
                    train[k] = eager.get(k)
                train["eager_variant"] = eager
            else:
                train["cuda_graph_variant"] = train_graph

    line = None
    if rank == 0:
        live_peaks = measure_dense_peaks(dev) if not args.no_extra else None
        roof, shares, prof_total_ms, per_launch_us = profile_step(pipe, batch_dev, peaks, args.dtype,
                                                                  live_peaks)
        roof["peak_source"] = peak_src
        # the event-bracketed launch duration above is measured in an eager step, where an event
        # record sits between the kernels and the prologue overlap of programmatic dependent launch
        # is lost (the eager kernel sum is eager_step_kernel_ms, the graph replay ms_per_step);
        # the kernel's time inside the timed graph replay, by its share of the step:
        if roof.get("share_of_step") and roof.get("avg_launch_us") and prof_total_ms:
            n_launch = roof["share_of_step"] * prof_total_ms * 1e3 / roof["avg_launch_us"]
            in_graph_us = roof["share_of_step"] * (dev_ms / args.steps) * 1e3 / max(1.0, n_launch)
            unit_work = roof["algorithmic_flops"] if roof["bound"] == "tensor" else roof["algorithmic_bytes"]
            scale = 1e12 if roof["bound"] == "tensor" else 1e9
            roof["in_graph"] = {
                "launch_us": in_graph_us, "launches_per_step": round(n_launch),
                "achieved": unit_work / (in_graph_us * 1e-6) / scale,
                "frac": unit_work / (in_graph_us * 1e-6) / scale / roof["peak"],
                "how": "share_of_step x graph-replay step time / launches of the kernel per step "
                       "(informational: `achieved` / `frac` above are the event-bracketed figures)"}
        sample_B = w["B"]
        fps_cpu, ms_cpu, cores = (cpu_oracle_arm(CPU_STEPS, CPU_WARMUP, sample_B)
                                  if world == 1 and not args.no_cpu else (None, None, None))
        eager = gpu_eager_baseline(dev) if world == 1 and not args.no_extra else None
        h2d = sum(t_.numel() * t_.element_size() for t_ in host)
        d2h = 4 + w["B"] * w["T"] * 8 + w["B"] * 4
        fpf = flops_per_frame()
        value = total_frames / (dev_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic",
            "config": {"workload": w["name"], "batch_per_gpu": w["B"], "T": w["T"],
                       "feat": w.get("feat", 256), "layers": 12, "vocab": w["vocab"],
                       "target_len": w["Lmax"], "what": w["desc"],
                       "valid_frames_per_step_per_gpu": frames_per_step,
                       "step": "encoder fwd" + ("" if w["kind"] == "single" else " + AV fusion")
                               + " + CTC loss + greedy decode",
                       "cuda_graph": not args.no_graph,
                       "l2": "256 MB memset between steps (outside the per-step event pairs)",
                       "timing": "CUDA events per step, summed over steps, max over ranks",
                       "parallelism": f"dp{world} (utterance-sharded, no data-path collective)"},
            "clocks": clocks,
            "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "EncoderCTCPipeline.run_stream (H2D of batch n+1 overlaps batch n; "
                           "L2 flush inside the timed region)",
                    "blocking_call_value": total_frames / (e2e_call_ms * 1e-3),
                    "blocking_call_api": "EncoderCTCPipeline.run, one blocking call per batch"},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "roofline": roof,
            "kernel_time_shares": shares,
            "kernel_us_per_launch": per_launch_us,
            "eager_step_kernel_ms": prof_total_ms,
            "other_mode": other_mode,
            "strong": strong,
            "train": train,
            "gpu_eager_baseline": eager,
            "model_tflops": value * fpf / 1e12,
            "model_tflops_per_gpu": value * fpf / 1e12 / world,
            "model_frac_of_bf16_sustained": value * fpf / 1e12 / world / peaks["bf16_tflops_sustained"],
            "loss": loss_ref,
        }
        if fps_cpu is not None:
            line["cpu_baseline"] = {
                "value": fps_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"the full {w['name']} batch per step ({sample_B}x{w['T']} frames), median of "
                          f"{CPU_STEPS} steps after {CPU_WARMUP} warm-up ({ms_cpu:.0f} ms/step), fp32 "
                          f"oracle port, torch {torch.__version__}"}
        else:
            line["cpu_baseline"] = None
        if args.mode == "train" and train and "value" in train:
            # the training step as the headline: same metric (valid frames/s), training semantics
            line["inference"] = {k: line[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches")}
            line.update(value=train["value"], ms_per_step=train["ms_per_step"], dtype="tf32",
                        steps=train["steps"], warmup=train["warmup"],
                        gpu_launches=train["kernel_launches_per_step"] * train["steps"])
            line["e2e"] = {"value": train["value"], "unit": UNIT, "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": 4,
                           "api": "encoder(...) / ctc(...) / loss.backward() / GradBucketReducer.finish()"}
            line["config"]["step"] = train["step"]
            line["config"]["mode"] = "train"
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of graph replay")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS),
                    help="SURVEY.md §8d workload (default: the bench line, C2)")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (sweep)")
    ap.add_argument("--T", type=int, default=0, help="encoder frames per utterance (sweep)")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the strong-scaling / training / eager-baseline legs and the live peaks")
    ap.add_argument("--train-steps", type=int, default=5, help="steps of the training leg")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: the training step becomes the headline value of the line")
    ap.add_argument("--dtype", default="bf16", choices=["tf32", "tf32x3", "bf16"],
                    help="compute mode (tailored_avsr_b200.engine): operand storage of the tensor-core products")
    args = ap.parse_args()
    select_workload(args)
    if args.impl == "reference":
        return run_reference_arm(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
