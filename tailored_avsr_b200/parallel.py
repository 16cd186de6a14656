"""Utterance-sharded data parallelism.

Every operation of the encoder + CTC path is per-utterance, so N GPUs simply take disjoint subsets
of the batch — there is no data-path collective.  The only cross-utterance term is the loss
normalisation `sum_b nll_b / B` (src/ctc/ctc.py:62-66): each rank divides by the GLOBAL batch size
and one scalar all-reduce (sum) recovers the reference value.  The training-mode gradient
all-reduce named in BASELINE.json needs the encoder backward kernels and is not built yet.

One process per GPU, launched by torch.distributed.run; NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_utterances(lens: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Indices of the utterances rank `rank` processes: utterances are sorted by length
    (longest first) and dealt round-robin, which balances the padded work across ranks and keeps
    the assignment deterministic.  The result is returned in increasing index order."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    order = sorted(range(len(lens)), key=lambda i: (-int(lens[i]), i))
    return sorted(order[rank::world_size])


def global_ctc_loss(local_nll: torch.Tensor, global_batch: int,
                    group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """sum_b nll_b / B over ALL ranks from each rank's per-utterance NLL vector."""
    part = local_nll.sum() / float(global_batch)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        part = part.clone()
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


def gather_token_lists(local_tokens: List[List[int]], local_indices: List[int], global_batch: int,
                       group: Optional[dist.ProcessGroup] = None) -> List[List[int]]:
    """Reassemble the greedy-decoded token lists of the whole batch on every rank (host side)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out: List[List[int]] = [[] for _ in range(global_batch)]
        for i, t in zip(local_indices, local_tokens):
            out[i] = t
        return out
    gathered: List[Optional[list]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, list(zip(local_indices, local_tokens)), group=group)
    out = [[] for _ in range(global_batch)]
    for part in gathered:
        for i, t in part:
            out[i] = t
    return out


def max_over_ranks(value: float, device: torch.device,
                   group: Optional[dist.ProcessGroup] = None) -> float:
    """Timing rule of bench.py: a multi-GPU step takes as long as its slowest rank."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


class GradBucketReducer:
    """Bucketed gradient all-reduce (the only collective of the path, training
    mode).  Gradients of `params` are packed into flat fp32 buckets of ~`bucket_mb` MB in REVERSE
    parameter order (the order backward produces them), each bucket is all-reduced (sum)
    asynchronously as soon as it is packed, and the results are copied back after the last launch,
    so the NCCL transfers of early buckets overlap the packing of later ones.  Ranks hold disjoint
    utterance shards and normalise their loss by the GLOBAL batch (`global_ctc_loss`), so the SUM
    of the per-rank gradients is the full-batch gradient: no division by the world size.

    Today the trainable part of the B200 path is the CTC head (`CTC.forward` has its CUDA backward;
    the encoder backward kernels are not built yet), so this runs on `ctc.parameters()`; it takes
    any parameter list.  gloo on CPU (tests) and NCCL on GPUs use the same code."""

    def __init__(self, params, bucket_mb: float = 25.0, group: Optional[dist.ProcessGroup] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        cap = max(1, int(bucket_mb * (1 << 20) // 4))
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self.buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)

    def reduce(self) -> int:
        """All-reduce (sum) every `.grad` in place; parameters without a gradient contribute zeros
        (every rank must launch the same collectives).  Returns the number of collectives issued."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return 0
        works = []
        for i, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            ref = bucket[0]
            flat = self._flat[i]
            if flat is None or flat.numel() != n or flat.device != ref.device:
                flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=ref.device)
            off = 0
            for p in bucket:
                view = flat[off:off + p.numel()]
                if p.grad is None:
                    view.zero_()
                else:
                    view.copy_(p.grad.reshape(-1))
                off += p.numel()
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for i, (bucket, work) in enumerate(zip(self.buckets, works)):
            work.wait()
            flat = self._flat[i]
            off = 0
            for p in bucket:
                g = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += p.numel()
        return len(works)
