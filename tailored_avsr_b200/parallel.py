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
