"""Utterance-sharded data parallelism.

Every operation of the encoder + CTC path is per-utterance, so N GPUs simply take disjoint subsets
of the batch — there is no data-path collective.  The only cross-utterance term is the loss
normalisation `sum_b nll_b / B` (src/ctc/ctc.py:62-66): each rank divides by the GLOBAL batch size
and one scalar all-reduce (sum) recovers the reference value.  The only collective of the path is
the training-mode gradient all-reduce (BASELINE.json north_star): `GradBucketReducer`, bucketed and
launched from gradient hooks WHILE the backward of the earlier blocks still runs.

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
    """Bucketed gradient all-reduce overlapped with the backward pass (the only
    collective of the path, training mode).

    The parameters are packed into flat fp32 buckets of ~`bucket_mb` MB in REVERSE registration
    order - the order the backward produces gradients: the CTC head first, then after_norm, block
    11, block 10, ... (training.py makes each block one autograd node, so a block's gradients all
    arrive when its node returns).  With `overlap=True` every parameter gets a
    post-accumulate-grad hook; when the last gradient of a bucket has arrived the bucket is packed
    and its `all_reduce(sum)` is launched asynchronously on the process group's stream, so the NCCL
    transfer of block k's bucket runs under the backward kernels of blocks < k.  Buckets are always
    launched IN ORDER (bucket i+1 never before bucket i): ranks may produce gradients in different
    patterns (stochastic depth skips different layers on different ranks), but the sequence of
    collectives must be identical everywhere.  `finish()` launches what is left (parameters without
    a gradient contribute zeros), waits, and copies the sums back into `.grad`.

    Ranks hold disjoint utterance shards and normalise their loss by the GLOBAL batch
    (`global_ctc_loss`), so the SUM of the per-rank gradients is the full-batch gradient: no
    division by the world size.  gloo on CPU (tests) and NCCL on GPUs use the same code."""

    def __init__(self, params, bucket_mb: float = 25.0, group: Optional[dist.ProcessGroup] = None,
                 overlap: bool = False):
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
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._works: List[Optional[object]] = [None] * len(self.buckets)
        self._pending = [len(b) for b in self.buckets]
        self._seen = set()
        self._next = 0
        self.launched_in_backward = 0
        self._hooks = []
        self.overlap = overlap
        if overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # ------------------------------------------------------------------------------------------
    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def bucket_bytes(self) -> List[int]:
        return [4 * sum(p.numel() for p in b) for b in self.buckets]

    def _launch(self, i: int) -> None:
        bucket = self.buckets[i]
        n = sum(p.numel() for p in bucket)
        ref = bucket[0]
        flat = self._flat[i]
        if flat is None or flat.numel() != n or flat.device != ref.device:
            flat = self._flat[i] = torch.empty(n, dtype=torch.float32, device=ref.device)
        # pack with ONE multi-tensor copy per bucket (a per-parameter copy_ is ~500 tiny launches
        # per step on this model); parameters without a gradient contribute zeros
        off = 0
        dst, src, missing = [], [], []
        for p in bucket:
            view = flat[off:off + p.numel()]
            if p.grad is None:
                missing.append(view)
            else:
                dst.append(view)
                src.append(p.grad.reshape(-1))
            off += p.numel()
        if dst:
            torch._foreach_copy_(dst, src)
        if missing:
            torch._foreach_zero_(missing)
        self._works[i] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        """Gradient hook: count the bucket down, launch every bucket that has become launchable."""
        if not self._active() or id(p) in self._seen:
            return
        self._seen.add(id(p))
        i = self._bucket_of[id(p)]
        self._pending[i] -= 1
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1
            self.launched_in_backward += 1

    def finish(self) -> int:
        """Launch the buckets that did not complete during backward, wait for all of them (the
        current stream waits; the host does not block on NCCL) and write the sums into `.grad`.
        Returns the number of collectives of this step."""
        if not self._active():
            return 0
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        for i, bucket in enumerate(self.buckets):
            self._works[i].wait()
            flat = self._flat[i]
            off = 0
            dst, src = [], []
            for p in bucket:
                g = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    dst.append(p.grad)
                    src.append(g)
                off += p.numel()
            if dst:
                torch._foreach_copy_(dst, src)
        n = len(self.buckets)
        self._works = [None] * n
        self._pending = [len(b) for b in self.buckets]
        self._seen = set()
        self._next = 0
        return n

    def reduce(self) -> int:
        """Non-overlapped form: call after backward() has returned."""
        self.launched_in_backward = 0
        return self.finish()

    def remove_hooks(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
