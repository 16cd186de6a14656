"""B200 drop-in for the reference `CTC` module (src/ctc/ctc.py:8-188): same constructor, `ctc_lo`
parameter names, `forward/softmax/log_softmax/argmax` and the `reduce` attribute; the head, the
log-domain forward-backward recursion and the greedy decode run as CUDA kernels."""
from __future__ import annotations

import logging
from typing import List, Optional

import torch
import torch.nn.functional as F

from .. import ops


class _CTCHeadLossFn(torch.autograd.Function):
    """nll_b = CTC(log_softmax(hs . W^T + b)) with every step on the CUDA path: the head and the
    forward-backward recursion in forward (the loss kernel also emits d nll_b / d logits =
    softmax - occupancy), the head's backward (d hs, d W, d b) in backward."""

    @staticmethod
    def forward(ctx, hs2d, weight, bias, targets, hlens, tlens, B, T, zero_infinity):
        logp, _, _ = ops.ctc_head(hs2d, weight, bias, want_logp=True)
        nll, grad = ops.ctc_loss(logp.view(B, T, -1), targets, hlens, tlens, want_grad=True,
                                 gscale=1.0, zero_infinity=zero_infinity)
        ctx.save_for_backward(grad, hs2d, weight)
        ctx.T = T
        return nll

    @staticmethod
    def backward(ctx, gout):
        grad, hs2d, weight = ctx.saved_tensors
        V = weight.shape[0]
        if V <= 64:
            dhs, dw, db = ops.ctc_head_bwd(grad, hs2d, weight.contiguous(),
                                           row_scale=gout.contiguous().float(), rows_per_seg=ctx.T)
        else:
            # larger vocabularies: scale d logits per utterance, then dgrad / wgrad on the tcgen05
            # GEMM like every other Linear of the training path
            if V % 4 != 0:
                raise NotImplementedError("CTC head backward for 64 < odim <= 256 needs odim % 4 == 0")
            from .. import ops_backward as ob
            g2 = grad.reshape(-1, V)
            zero = torch.zeros_like(gout, dtype=torch.float32)
            dl = ops.scale_add_rows(g2, g2, gout.contiguous().float(), zero, ctx.T)
            dhs, dw, db = ob.linear_bwd(hs2d, weight.contiguous(), dl)
        return dhs, dw, db, None, None, None, None, None, None


class CTC(torch.nn.Module):
    """CTC module (B200 path).

    Args (ctc.py:21-30): odim, encoder_output_size, dropout_rate, ctc_type ("builtin"), reduce,
    ignore_nan_grad, zero_infinity.
    """

    def __init__(self, odim: int, encoder_output_size: int, dropout_rate: float = 0.0,
                 ctc_type: str = "builtin", reduce: bool = True, ignore_nan_grad: bool = None,
                 zero_infinity: bool = True):
        super().__init__()
        eprojs = encoder_output_size
        self.dropout_rate = dropout_rate
        self.ctc_lo = torch.nn.Linear(eprojs, odim)
        self.ctc_type = ctc_type
        if ignore_nan_grad is not None:
            zero_infinity = ignore_nan_grad
        if self.ctc_type == "builtin":
            self.zero_infinity = bool(zero_infinity)
        elif self.ctc_type in ("builtin2", "gtnctc"):
            raise NotImplementedError(f'ctc_type="{self.ctc_type}" is not built on the B200 path '
                                      '(every shipped config uses "builtin")')
        else:
            raise ValueError(f'ctc_type must be "builtin" or "gtnctc": {self.ctc_type}')
        self.reduce = reduce
        if odim > 256:
            raise NotImplementedError(
                f"tailored_avsr_b200.CTC: the CUDA head / loss kernels are built for vocabularies up "
                f"to 256 (char-level EN 41 / ES 37 and the 256-token SentencePiece alternative); "
                f"odim={odim}")

    # ---------------------------------------------------------------------------------------
    def _head(self, hs_pad: torch.Tensor, logp=False, prob=False, amax=False):
        if not hs_pad.is_cuda:
            raise RuntimeError("tailored_avsr_b200.CTC runs on CUDA tensors only (no CPU fallback)")
        B, T, D = hs_pad.shape
        hs2 = hs_pad.reshape(B * T, D).contiguous().float()
        return ops.ctc_head(hs2, self.ctc_lo.weight, self.ctc_lo.bias, logp, prob, amax), B, T

    def forward(self, hs_pad, hlens, ys_pad, ys_lens):
        """CTC loss (ctc.py:133-158): sum_b nll_b / B, or the vector nll_b / B if not self.reduce.

        The reference applies F.dropout with the functional default training=True even in eval
        (ctc.py:143); that call is kept on the host side so the RNG stream is the reference's.
        """
        hs = F.dropout(hs_pad, p=self.dropout_rate) if self.dropout_rate > 0 else hs_pad
        B, T, _ = hs.shape
        dev = hs.device
        targets = ys_pad.to(dev).long().contiguous()
        hl = hlens.to(dev).to(torch.int32)
        tl = ys_lens.to(dev).to(torch.int32)
        needs_grad = torch.is_grad_enabled() and (hs.requires_grad or self.ctc_lo.weight.requires_grad)
        if needs_grad:
            # training: head + loss + their backward all on the CUDA path (no ATen arithmetic)
            D = hs.shape[-1]
            hs2d = hs.reshape(B * T, D).contiguous().float()
            nll = _CTCHeadLossFn.apply(hs2d, self.ctc_lo.weight, self.ctc_lo.bias, targets, hl, tl,
                                       B, T, self.zero_infinity)
        else:
            (logp, _, _), _, _ = self._head(hs, logp=True)
            nll, _ = ops.ctc_loss(logp.view(B, T, -1), targets, hl, tl, zero_infinity=self.zero_infinity)
        loss = nll.sum() / B if self.reduce else nll / B
        return loss.to(device=hs_pad.device, dtype=hs_pad.dtype)

    def softmax(self, hs_pad):
        (_, prob, _), B, T = self._head(hs_pad, prob=True)
        return prob.view(B, T, -1)

    def log_softmax(self, hs_pad):
        (logp, _, _), B, T = self._head(hs_pad, logp=True)
        return logp.view(B, T, -1)

    def argmax(self, hs_pad):
        (_, _, amax), B, T = self._head(hs_pad, amax=True)
        return amax.view(B, T)

    def loss_and_greedy(self, hs_pad, hlens, ys_pad, ys_lens, greedy_lens=None, blank: int = 0):
        """Validation-step fusion (espnet_model.py:586-592): ONE head launch yields the
        log-softmax for the loss and the argmax for the greedy decode.  Returns
        (loss, tokens (B,T) padded with -1, ntok (B,)).  Inference only (dropout_rate must be 0
        or the module in the reference's always-on dropout mode is bypassed: use forward())."""
        if self.dropout_rate > 0:
            loss = self.forward(hs_pad, hlens, ys_pad, ys_lens)
            tokens, ntok = self.greedy(hs_pad, greedy_lens, blank)
            return loss, tokens, ntok
        B, T, _ = hs_pad.shape
        dev = hs_pad.device
        (logp, _, amax), _, _ = self._head(hs_pad, logp=True, amax=True)
        nll, _ = ops.ctc_loss(logp.view(B, T, -1), ys_pad.to(dev).long().contiguous(),
                              hlens.to(dev).to(torch.int32), ys_lens.to(dev).to(torch.int32),
                              zero_infinity=self.zero_infinity)
        loss = (nll.sum() / B if self.reduce else nll / B).to(dtype=hs_pad.dtype)
        lens = None if greedy_lens is None else greedy_lens.to(dev).to(torch.int32)
        tokens, ntok = ops.ctc_greedy(amax.view(B, T), lens, blank)
        return loss, tokens, ntok

    # ---- extras of the B200 path (device-side replacement of the host groupby loops) ----
    def greedy(self, hs_pad, hlens: Optional[torch.Tensor] = None, blank: int = 0):
        """argmax -> collapse repeats -> drop blank on the device (espnet_model.py:590-592,
        maskctc_model.py:287-291).  Returns (tokens (B,T) padded with -1, ntok (B,))."""
        amax = self.argmax(hs_pad)
        lens = None if hlens is None else hlens.to(hs_pad.device).to(torch.int32)
        return ops.ctc_greedy(amax, lens, blank)

    def greedy_lists(self, hs_pad, hlens=None, blank: int = 0) -> List[List[int]]:
        tokens, ntok = self.greedy(hs_pad, hlens, blank)
        tokens, ntok = tokens.cpu(), ntok.cpu()
        return [tokens[b, : int(ntok[b])].tolist() for b in range(tokens.shape[0])]
