"""Batched CTC beam search on the device-resident prefix scorer - the decoding loop the reference
runs through espnet's `BatchBeamSearch` with the `ctc` partial scorer
(src/inference/asr_inference.py:142 builds the scorer, :276-303 the search object, :474 calls it).

espnet's loop, per output position i (BatchBeamSearch.search / .post_process):
    scores[h, c]  = partial scorer (CTC prefix score of h.c minus that of h)
    weighted      = hyp.score[h] + ctc_weight * scores[h, c] (+ length_bonus per emitted token)
    top-`beam` of the flattened (hyp, token) table -> new hypotheses (prev hyp id, new token id)
    hypotheses ending in <eos> leave the running set and join the ended list
Here every one of those steps is a device op over ALL running hypotheses at once, the scorer state
(r (T,2) forward variables and log psi per hypothesis) never leaves the GPU, shapes are static
(the running set always holds `beam` slots, dead slots carry -inf), and the host only reads ONE
flag every `check_every` steps to stop early - no per-step host synchronisation on scores.

Stopping rule: with a pure CTC score (length_bonus = 0) a hypothesis' score can only fall when it
is extended, so once the best ended hypothesis scores at least as high as the best running one the
n-best list is final; otherwise the loop runs to `maxlen` (= T frames, espnet's maxlenratio = 0).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from .prefix_scorer import CTCPrefixScorer

NEG = -1e30


class CTCBeamSearch:
    def __init__(self, ctc: torch.nn.Module, beam_size: int, sos: int, eos: int,
                 ctc_weight: float = 1.0, length_bonus: float = 0.0, check_every: int = 4):
        self.scorer = CTCPrefixScorer(ctc=ctc, eos=eos)
        self.beam = beam_size
        self.sos, self.eos = sos, eos
        self.w = ctc_weight
        self.bonus = length_bonus
        self.check_every = max(1, check_every)

    @torch.no_grad()
    def search(self, x: torch.Tensor, maxlen: Optional[int] = None, nbest: int = 1
               ) -> List[Tuple[List[int], float]]:
        """x (T, D): encoder output of ONE utterance on the device.  Returns the n-best list
        [(token ids without sos / eos, score)], best first."""
        sc = self.scorer
        sc.batch_init_state(x)
        dev = x.device
        T, V, K = sc.T, sc.logp.shape[1], self.beam
        maxlen = T if maxlen is None else maxlen
        # running set: K slots (slot 0 live at the start)
        y = torch.full((1, 1), self.sos, dtype=torch.int64, device=dev)
        hyp_score = torch.zeros((1,), device=dev)
        states = [None]
        # ended pool: the best K finished hypotheses so far, padded token matrix
        end_y = torch.full((K, maxlen + 2), -1, dtype=torch.int64, device=dev)
        end_score = torch.full((K,), NEG, device=dev)
        end_len = torch.zeros((K,), dtype=torch.int64, device=dev)
        for i in range(maxlen):
            score, (r_new, psi_new) = sc.batch_score_partial(y, None, states, x)
            total = hyp_score[:, None] + self.w * score + self.bonus          # (n, V)
            total[:, sc.blank] = NEG
            if i == maxlen - 1:       # last position: only <eos> may follow (espnet post_process)
                keep = total[:, self.eos].clone()
                total.fill_(NEG)
                total[:, self.eos] = keep
            n = total.shape[0]
            k = min(K, n * V)
            top, idx = total.reshape(-1).topk(k)
            prev, tok = idx // V, idx % V
            if k < K:                 # first step: fewer candidates than slots -> dead padding
                pad = K - k
                top = torch.cat([top, torch.full((pad,), NEG, device=dev)])
                prev = torch.cat([prev, torch.zeros((pad,), dtype=torch.int64, device=dev)])
                tok = torch.cat([tok, torch.full((pad,), self.eos, dtype=torch.int64, device=dev)])
            new_y = torch.cat([y[prev], tok[:, None]], 1)                      # (K, i + 2)
            ended = tok == self.eos
            # ---- ended hypotheses join the pool (keep its best K) ----
            cand_score = torch.where(ended, top, torch.full_like(top, NEG))
            cand_y = torch.full((K, maxlen + 2), -1, dtype=torch.int64, device=dev)
            cand_y[:, : i + 2] = new_y
            pool_s = torch.cat([end_score, cand_score])
            pool_y = torch.cat([end_y, cand_y])
            pool_l = torch.cat([end_len, torch.full((K,), i + 2, dtype=torch.int64, device=dev)])
            end_score, sel = pool_s.topk(K)
            end_y, end_len = pool_y[sel], pool_l[sel]
            # ---- the others keep running (dead slots: -inf) ----
            hyp_score = torch.where(ended, torch.full_like(top, NEG), top)
            y = new_y
            r_sel = r_new[prev, :, tok, :]                                     # (K, T, 2)
            psi_sel = psi_new[prev, tok]
            states = [(r_sel[j], psi_sel[j]) for j in range(K)]
            if (i + 1) % self.check_every == 0 or i == maxlen - 1:
                best_run = hyp_score.max()
                done = (best_run <= NEG / 2) | ((self.bonus <= 0) & (end_score[nbest - 1] >= best_run)
                                                & (end_score[nbest - 1] > NEG / 2))
                if bool(done):        # the ONLY host read of the loop: one flag
                    break
        out = []
        es, el, ey = end_score.cpu(), end_len.cpu(), end_y.cpu()
        for j in range(min(nbest, K)):
            if float(es[j]) <= NEG / 2:
                break
            toks = ey[j, 1: int(el[j]) - 1].tolist()
            out.append((toks, float(es[j])))
        return out
