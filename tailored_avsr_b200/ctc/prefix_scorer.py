"""Device-resident CTC prefix scorer for beam search — the object the reference builds at
`src/inference/asr_inference.py:142` (`CTCPrefixScorer(ctc=asr_model.ctc, eos=asr_model.eos)`,
espnet `BatchPartialScorerInterface` over `CTCPrefixScoreTH`).

Same method surface as the espnet scorer (`init_state`, `batch_init_state`, `score_partial`,
`batch_score_partial`, `select_state`), but every decoding step is ONE launch of
`tavsr_ctc_prefix_score` over all live hypotheses x all V tokens x all T frames, with the forward
variables kept on the device between steps — instead of espnet's Python `for t in range(start, T)`
loop of tiny kernels per step.

State of a hypothesis h: `(r (T,2) = (log r^n, log r^b), log_psi = log-prob of h as a prefix)`.
Scores returned are `log psi(h.c) - log psi(h)` like the reference's; `eos` scores the complete
sequence, `blank` scores logzero (-1e10).
"""
from typing import Any, List, Optional, Tuple

import torch

from .. import ops


class CTCPrefixScorer:
    def __init__(self, ctc: torch.nn.Module, eos: int):
        self.ctc = ctc
        self.eos = eos
        self.blank = 0
        self.logp: Optional[torch.Tensor] = None  # (T, V) log-softmax of the utterance
        self.T = 0

    # ---- utterance setup -------------------------------------------------------------------
    def _setup(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if not x.is_cuda:
            raise RuntimeError("tailored_avsr_b200.CTCPrefixScorer runs on CUDA tensors only "
                               "(no CPU fallback)")
        with torch.no_grad():
            logp = self.ctc.log_softmax(x.unsqueeze(0))[0].contiguous()  # batch_init_state: A.9
        self.logp = logp
        self.T = logp.shape[0]
        r0 = torch.full((self.T, 2), -1e10, device=x.device, dtype=torch.float32)
        r0[:, 1] = torch.cumsum(logp[:, self.blank], dim=0)
        self._init = (r0, torch.zeros((), device=x.device, dtype=torch.float32))
        return self._init

    def init_state(self, x: torch.Tensor):
        """State of the empty prefix for encoder output x (T, D)."""
        return self._setup(x)

    def batch_init_state(self, x: torch.Tensor):
        """espnet returns None here and creates the first state lazily; so do we."""
        self._setup(x)
        return None

    # ---- one decoding step -----------------------------------------------------------------
    def batch_score_partial(self, y: torch.Tensor, ids: Optional[torch.Tensor], state: List[Any],
                            x: torch.Tensor):
        """y (n, ylen) int64 prefixes starting with sos; state: list of n selected states (or
        None entries before the first step).  Returns (scores (n, V), batched new state)."""
        n, ylen = y.shape
        dev = self.logp.device
        sel = [s if s is not None else self._init for s in state]
        r_prev = torch.stack([s[0] for s in sel]).contiguous()
        psi_prev = torch.stack([s[1] for s in sel]).contiguous()
        plen = torch.full((n,), ylen - 1, device=dev, dtype=torch.int32)
        if ylen > 1:
            last = y[:, -1].to(dev).to(torch.int32).contiguous()
        else:
            last = torch.full((n,), -1, device=dev, dtype=torch.int32)
        r_new, score = ops.ctc_prefix_score(self.logp, r_prev, last, plen, psi_prev, self.T,
                                            self.blank, self.eos)
        psi_new = score + psi_prev[:, None]
        if ids is not None:
            # partial scoring (pre-beam): tokens outside `ids` are not candidates
            keep = torch.zeros_like(score, dtype=torch.bool)
            keep.scatter_(1, ids.to(dev).long(), True)
            score = torch.where(keep, score, torch.full_like(score, -1e10))
        return score, (r_new, psi_new)

    def score_partial(self, y: torch.Tensor, ids: Optional[torch.Tensor], state: Any,
                      x: torch.Tensor):
        """Single-hypothesis form used by espnet's non-batch `BeamSearch` (the path
        asr_inference.py:276-303 falls back to): y (ylen,), state from init_state / select_state.
        Like espnet's scorer it returns ONE SCORE PER ENTRY OF `ids` (not a full-V vector -
        BeamSearch does `weighted_scores[part_ids] += w * part_scores`) and a state indexed by the
        position in `ids`: (r (len(ids), T, 2), log_psi (len(ids),))."""
        score, (r_new, psi_new) = self.batch_score_partial(y.unsqueeze(0), None, [state], x)
        if ids is None:
            ids = torch.arange(score.shape[1], device=score.device)
        idx = ids.to(score.device).long()
        return score[0, idx], (r_new[0][:, idx, :].permute(1, 0, 2), psi_new[0, idx])

    def select_state(self, state, i: int, new_id: Optional[int] = None):
        """Three-argument form (BatchBeamSearch): state = what batch_score_partial returned,
        i = hypothesis, new_id = the token it is extended by.  Two-argument form (BeamSearch):
        state = what score_partial returned, i = position of the chosen token in its `ids`.
        Both return the (r (T,2), log_psi) state of the extended hypothesis (views into the
        step's device buffers)."""
        if state is None:
            return None
        r_new, psi_new = state
        if new_id is None:
            return r_new[i], psi_new[i]
        return r_new[i, :, new_id, :], psi_new[i, new_id]
