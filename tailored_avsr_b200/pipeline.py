"""Encoder + CTC pipeline: the public "one call per batch" entry of the B200 path.

`EncoderCTCPipeline.run(feats, feats_lens, ys_pad, ys_lens)` does what a validation step of the
reference does with its encoder and CTC modules (src/models/espnet_model.py:397-402, 578-593):
encoder forward, CTC loss, greedy CTC decode.  Inputs may live on the host (they are staged through
pinned memory and copied on the compute stream) or on the device.

The ~160 kernel launches of a 12-block forward are captured ONCE per input-shape signature into a
CUDA graph (static input / output buffers, private memory pool) and replayed, which removes the
Python / launch overhead that otherwise dominates at B*T ~ 8k frames.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Sequence, Tuple

import torch


class EncoderCTCPipeline:
    def __init__(self, encoder: torch.nn.Module, ctc: torch.nn.Module, use_cuda_graph: bool = True,
                 greedy: bool = True):
        self.encoder = encoder.eval()
        self.ctc = ctc.eval()
        self.use_cuda_graph = use_cuda_graph
        self.greedy = greedy
        self._graphs: Dict[Tuple, dict] = {}
        self._param_sig = None
        p = next(encoder.parameters())
        if not p.is_cuda:
            raise RuntimeError("EncoderCTCPipeline needs the modules on a CUDA device "
                               "(there is no CPU fallback)")
        self.device = p.device

    # ---------------------------------------------------------------------------------------
    def _modules(self):
        return [m for m in (self.encoder, self.ctc, getattr(self, "fusion", None),
                            getattr(self, "acoustic_embed", None), getattr(self, "visual_embed", None))
                if m is not None]

    def _parameter_signature(self) -> tuple:
        """In-place version counters of every parameter and buffer the captured graphs read.  A
        CUDA graph bakes in derived tensors (fused QKV weight, folded merge weights, permuted conv
        weights) and host scalars taken from the parameters at capture time, so a replay after
        `load_state_dict` / an optimizer step would mix new raw parameters with stale derived
        ones: the signature is checked before every replay (one attribute read per tensor, ~30 us)
        and the graphs are re-captured when it changed.  Writes through `.data`
        (`p.data.copy_()`) do not bump the version counter and replacing parameter OBJECTS
        (`module.to()`, `load_state_dict(assign=True)`) is not seen either - call `invalidate()`
        after those."""
        tens = self.__dict__.get("_sig_tensors")
        if tens is None:
            tens = []
            for m in self._modules():
                tens += list(m.parameters()) + list(m.buffers())
            self._sig_tensors = tens
            self._sig_ptrs = tuple(t.data_ptr() for t in tens)
        try:
            return (self._sig_ptrs,) + tuple([t._version for t in tens])
        except RuntimeError:   # inference tensors do not track versions
            return (self._sig_ptrs,)

    def invalidate(self) -> None:
        """Drop every captured graph and derived-weight cache (call after writing parameters
        through `.data` or any other path that does not bump the tensors' version counters)."""
        from .engine import PackedCache
        self._graphs.clear()
        self._param_sig = None
        self.__dict__.pop("_sig_tensors", None)
        for m in self._modules():
            for sub in m.modules():
                pk = getattr(sub, "_packed", None)
                if isinstance(pk, PackedCache):
                    pk.clear()

    def _step(self, feats, feats_lens, ys_pad, ys_lens):
        out, olens, _ = self.encoder(feats, feats_lens)
        if isinstance(out, tuple):
            out = out[0]
        if self.greedy and hasattr(self.ctc, "loss_and_greedy"):
            # like _calc_ctc_loss the collapse runs over all Tmax frames (espnet_model.py:590-592)
            loss, tokens, ntok = self.ctc.loss_and_greedy(out, olens, ys_pad, ys_lens)
            return {"encoder_out": out, "olens": olens, "loss": loss, "tokens": tokens, "ntok": ntok}
        loss = self.ctc(out, olens, ys_pad, ys_lens)
        res = {"encoder_out": out, "olens": olens, "loss": loss}
        if self.greedy:
            res["tokens"], res["ntok"] = self.ctc.greedy(out)
        return res

    def _capture(self, key, *tensors):
        static_in = [torch.empty_like(t) for t in tensors]
        for s_, t in zip(static_in, tensors):
            s_.copy_(t)
        # warm-up on a side stream: fills the weight-pack caches, the TMA descriptor cache and
        # the function attributes before capture
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                self._step(*static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            static_out = self._step(*static_in)
        entry = {"graph": graph, "in": static_in, "out": static_out}
        self._graphs[key] = entry
        return entry

    @torch.no_grad()
    def run_device(self, *tensors) -> dict:
        """Device tensors in, device tensors out (results alias static buffers when graphs are on:
        consume or clone them before the next call)."""
        if not self.use_cuda_graph:
            return self._step(*tensors)
        from .engine import compute_dtype
        key = tuple((tuple(t.shape), t.dtype) for t in tensors) + (compute_dtype(),)
        sig = self._parameter_signature()
        if sig != self._param_sig:
            self._graphs.clear()       # parameters changed since capture: derived tensors are stale
            self._param_sig = sig
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(key, *tensors)
        for s_, t in zip(entry["in"], tensors):
            s_.copy_(t, non_blocking=True)
        entry["graph"].replay()
        return entry["out"]

    def replay_static(self, key=None) -> dict:
        """Replay the captured graph on whatever currently sits in its static input buffers
        (bench.py's device-resident timing)."""
        entry = next(iter(self._graphs.values())) if key is None else self._graphs[key]
        entry["graph"].replay()
        return entry["out"]

    @torch.no_grad()
    def run_stream(self, batches: Iterable[Sequence[torch.Tensor]]) -> Iterator[dict]:
        """Pipelined form of `run` for a stream of host batches (pinned memory): the host->device
        copy of batch n+1 is issued on a copy stream before batch n is computed, so PCIe transfers
        hide under the kernels.  Yields, in order, the same host-side dict as `run` (loss, tokens,
        ntok on the host) plus `olens`; every batch still crosses PCIe and every result is read
        back before the next one is yielded."""
        dev = self.device
        comp = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy = self._copy_stream

        stages = self.__dict__.setdefault("_stages", {})
        turn = [0]

        def stage(batch):
            # two persistent device staging sets per shape signature: no allocator traffic on the
            # copy stream, and the set batch n+1 lands in is not the one batch n is read from
            key = tuple((tuple(t.shape), t.dtype) for t in batch)
            sets = stages.get(key)
            if sets is None:
                sets = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in batch]
                        for _ in range(2)]
                stages[key] = sets
            tens = sets[turn[0] & 1]
            turn[0] += 1
            copy.wait_stream(comp)  # the batch that last used this set has been consumed
            with torch.cuda.stream(copy):
                for d, t in zip(tens, batch):
                    d.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return tens, ev

        # results leave the device asynchronously too: right after the replay of batch n its loss /
        # tokens are copied (stream-ordered, so before batch n+1 overwrites the graph's static
        # outputs) into one of two pinned host sets, and batch n is only waited for AFTER batch n+1
        # has been enqueued - the GPU never idles on the host reading a result back
        pinned = self.__dict__.setdefault("_result_pins", {})

        def read_back(res, slot):
            keys = ["loss"] + (["tokens", "ntok"] if self.greedy else [])
            sig = tuple((k, tuple(res[k].shape), res[k].dtype) for k in keys) + (slot,)
            bufs = pinned.get(sig)
            if bufs is None:
                bufs = pinned[sig] = {k: torch.empty(res[k].shape, dtype=res[k].dtype).pin_memory()
                                      for k in keys}
            for k in keys:
                bufs[k].copy_(res[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(comp)
            return bufs, ev, res["olens"]

        it = iter(batches)
        try:
            nxt = stage(next(it))
        except StopIteration:
            return
        pending = None
        n = 0
        while nxt is not None:
            cur = nxt
            try:
                nxt = stage(next(it))
            except StopIteration:
                nxt = None
            comp.wait_event(cur[1])
            res = self.run_device(*cur[0])
            done = read_back(res, n & 1)
            n += 1
            if pending is not None:
                bufs, ev, olens = pending
                ev.synchronize()
                yield dict({k: v.clone() for k, v in bufs.items()}, olens=olens)
            pending = done
        if pending is not None:
            bufs, ev, olens = pending
            ev.synchronize()
            yield dict({k: v.clone() for k, v in bufs.items()}, olens=olens)

    @torch.no_grad()
    def run(self, *tensors) -> dict:
        """Host or device tensors in (feats, feats_lens, ys_pad, ys_lens for the single-stream
        pipeline); returns host-side loss (float tensor), token lists lengths and the device
        encoder output.  Host inputs should be pinned for an asynchronous copy."""
        dev = self.device
        res = self.run_device(*[t.to(dev, non_blocking=True) for t in tensors])
        out = {"encoder_out": res["encoder_out"], "olens": res["olens"],
               "loss": res["loss"].to("cpu", non_blocking=False)}
        if self.greedy:
            out["tokens"] = res["tokens"].to("cpu")
            out["ntok"] = res["ntok"].to("cpu")
        return out


class AVEncoderCTCPipeline(EncoderCTCPipeline):
    """Audio-visual form: `run(audio, video, lens_audio, lens_video, ys_pad, ys_lens)` with the two
    time-aligned (B, T, d) streams as they enter the encoder blocks (post-embed, x sqrt(d) applied:
    src/models/avsr_espnet_model.py:427-448).  One step = TailoredEncoder / ConventionalEncoder ->
    AdaptiveAudioVisualFusion -> CTC loss + greedy decode (avsr_espnet_model.py:451-467, 678-683).
    The key-padding masks and the relative positional table are built on the device."""

    def __init__(self, encoder, fusion, ctc, use_cuda_graph: bool = True, greedy: bool = True,
                 acoustic_embed=None, visual_embed=None, ignore_id: float = -1.0):
        """With `acoustic_embed` / `visual_embed` (DefaultEmbeddingLayerForAVSR) the inputs are the
        raw per-modality features and their lengths, and the step starts where the model's encode()
        does: embed -> temporal alignment (pad with ignore_id) -> positional encoding
        (avsr_espnet_model.py:427-448)."""
        super().__init__(encoder, ctc, use_cuda_graph=use_cuda_graph, greedy=greedy)
        from .espnet_compat import RelPositionalEncoding
        self.fusion = fusion.eval()
        self._pos = RelPositionalEncoding(encoder.output_size(), 0.0)
        assert (acoustic_embed is None) == (visual_embed is None), "pass both embeds or neither"
        self.acoustic_embed = acoustic_embed.eval() if acoustic_embed is not None else None
        self.visual_embed = visual_embed.eval() if visual_embed is not None else None
        self.ignore_id = ignore_id

    def _embed(self, audio, video, lens_a, lens_v):
        """Raw features -> aligned, position-encoded block inputs and masks."""
        import torch.nn.functional as F
        a, ma = self.acoustic_embed.apply_embed_layer(audio, lens_a)
        v, mv = self.visual_embed.apply_embed_layer(video, lens_v)
        pad = a.shape[1] - v.shape[1]   # static per input shape: the CUDA graph key covers it
        if pad < 0:
            a = F.pad(a, (0, 0, 0, -pad, 0, 0), value=self.ignore_id)
            ma = F.pad(ma, (0, -pad), value=False)
        elif pad > 0:
            v = F.pad(v, (0, 0, 0, pad, 0, 0), value=self.ignore_id)
            mv = F.pad(mv, (0, pad), value=False)
        (a, pos), (v, _) = self.acoustic_embed.apply_pos_enc(a), self.visual_embed.apply_pos_enc(v)
        return a, v, ma, mv, pos

    def _step(self, audio, video, lens_a, lens_v, ys_pad, ys_lens):
        dev = audio.device
        if self.acoustic_embed is not None:
            audio, video, mask_a, mask_v, pos = self._embed(audio, video, lens_a, lens_v)
            B, T, _ = audio.shape
        else:
            B, T, _ = audio.shape
            ar = torch.arange(T, device=dev)[None, :]
            mask_a = (ar < lens_a[:, None]).unsqueeze(1)
            mask_v = (ar < lens_v[:, None]).unsqueeze(1)
            pos = self._pos.pos_emb(T, dev)
        ya, _, yv, _, _ = self.encoder((audio, pos), mask_a, (video, pos), mask_v, ctc=self.ctc,
                                       audiovisual_fusion=self.fusion)
        if isinstance(ya, tuple):
            ya = ya[0]
        fused, olens = self.fusion(ya, mask_a, yv, mask_v)
        if self.greedy:
            loss, tokens, ntok = self.ctc.loss_and_greedy(fused, olens, ys_pad, ys_lens)
            return {"encoder_out": fused, "olens": olens, "loss": loss, "tokens": tokens, "ntok": ntok}
        return {"encoder_out": fused, "olens": olens,
                "loss": self.ctc(fused, olens, ys_pad, ys_lens)}
