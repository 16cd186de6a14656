"""world_size-2 gloo tests (CPU) of the utterance-sharded data-parallel host logic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_path, synth
from tailored_avsr_b200 import parallel


def test_shard_utterances_partition_and_balance():
    lens = [250, 100, 240, 90, 230, 80, 220, 70, 33]
    for world in (1, 2, 3, 4, 8):
        shards = [parallel.shard_utterances(lens, world, r) for r in range(world)]
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lens)))           # every utterance exactly once
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        work = [sum(lens[i] for i in s) for s in shards]
        assert max(work) - min(work) <= max(lens)        # length-balanced


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, T, V, L = 6, 40, 11, 9
        lens = torch.tensor([40, 25, 38, 31, 40, 12])
        hs = synth.randn((B, T, 16), 5)
        sd = {"ctc_lo.weight": synth.randn((V, 16), 6) * 0.3, "ctc_lo.bias": synth.randn((V,), 7) * 0.1}
        ys = synth.rand_targets(B, L, V, 8)
        ylens = torch.tensor([9, 5, 7, 6, 9, 9])           # last one infeasible for 12 frames? (no)
        mine = parallel.shard_utterances(lens.tolist(), world, rank)
        idx = torch.tensor(mine)
        nll_vec = ref_path.ctc_loss(hs[idx], lens[idx], ys[idx], ylens[idx], sd, reduce=False) * len(mine)
        loss = parallel.global_ctc_loss(nll_vec, B)
        toks = ref_path.ctc_greedy(hs[idx], sd, lens=lens[idx])
        all_toks = parallel.gather_token_lists(toks, mine, B)
        slowest = parallel.max_over_ranks(float(rank + 1), torch.device("cpu"))
        if rank == 0:
            full = ref_path.ctc_loss(hs, lens, ys, ylens, sd, reduce=True)
            ret["loss_ok"] = bool(torch.allclose(loss, full, rtol=1e-6, atol=1e-6))
            ret["toks_ok"] = all_toks == ref_path.ctc_greedy(hs, sd, lens=lens)
            ret["max_ok"] = slowest == float(world)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_loss_and_gather_match_single_process():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret.get("loss_ok") and ret.get("toks_ok") and ret.get("max_ok"), dict(ret)


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, T, D, V, L = 6, 30, 16, 11, 7
        lens = torch.tensor([30, 25, 28, 21, 30, 12])
        hs = synth.randn((B, T, D), 5)
        ys = synth.rand_targets(B, L, V, 8)
        ylens = torch.tensor([7, 5, 7, 6, 7, 3])
        torch.manual_seed(0)
        lo = torch.nn.Linear(D, V)
        extra = torch.nn.Parameter(torch.zeros(3))      # never used: contributes zeros

        def loss_of(idx):
            sd = {"ctc_lo.weight": lo.weight, "ctc_lo.bias": lo.bias}
            vec = ref_path.ctc_loss(hs[idx], lens[idx], ys[idx], ylens[idx], sd, reduce=False) * len(idx)
            return vec.sum() / B                          # normalised by the GLOBAL batch

        mine = torch.tensor(parallel.shard_utterances(lens.tolist(), world, rank))
        loss_of(mine).backward()
        # tiny buckets: weight (176 floats) and bias / extra land in different collectives
        red = parallel.GradBucketReducer([lo.weight, lo.bias, extra], bucket_mb=0.0005)
        n = red.reduce()
        got_w, got_b = lo.weight.grad.clone(), lo.bias.grad.clone()
        lo.zero_grad()
        loss_of(torch.arange(B)).backward()
        if rank == 0:
            ret["n_collectives"] = n
            ret["w_ok"] = bool(torch.allclose(got_w, lo.weight.grad, rtol=1e-5, atol=1e-6))
            ret["b_ok"] = bool(torch.allclose(got_b, lo.bias.grad, rtol=1e-5, atol=1e-6))
            ret["extra_ok"] = bool(extra.grad is not None and float(extra.grad.abs().sum()) == 0.0)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_bucketed_gradient_allreduce_equals_full_batch_gradient():
    """Sum of the per-rank gradients (loss normalised by the global batch) == full-batch gradient;
    buckets are formed in reverse parameter order and every rank launches the same collectives."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret.get("w_ok") and ret.get("b_ok") and ret.get("extra_ok"), dict(ret)
        assert ret.get("n_collectives") >= 2, dict(ret)


def test_grad_bucket_reducer_buckets_in_reverse_parameter_order():
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 300, 5, 200)]
    frozen = torch.nn.Parameter(torch.zeros(7), requires_grad=False)
    r = parallel.GradBucketReducer(ps + [frozen], bucket_mb=300 * 4 / (1 << 20))
    assert [[p.numel() for p in b] for b in r.buckets] == [[200, 5], [300], [10]]
    assert r.reduce() == 0          # no process group: nothing to do, gradients untouched


def _overlap_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # a 4-"block" chain; rank 1 skips block 2 (stochastic depth on one rank only), so its
        # gradient pattern differs from rank 0's: the collectives must still pair up in order
        blocks = [torch.nn.Linear(8, 8) for _ in range(4)]
        params = [p for b in blocks for p in b.parameters()]
        red = parallel.GradBucketReducer(params, bucket_mb=72 * 4 / (1 << 20), overlap=True)
        assert len(red.buckets) == 4
        x = synth.randn((5, 8), 3 + rank)

        def fwd(skip):
            h = x
            for i, b in enumerate(blocks):
                if i in skip:
                    continue
                h = torch.tanh(b(h))
            return h.sum()

        for step in range(2):                       # twice: the per-step state must reset
            for p in params:
                p.grad = None
            fwd({2} if rank == 1 else set()).backward()
            launched = red.launched_in_backward
            n = red.finish()
            got = [p.grad.clone() for p in params]
            # reference: both ranks' gradients computed locally and summed
            want = [torch.zeros_like(p) for p in params]
            for r in range(world):
                xr = synth.randn((5, 8), 3 + r)
                h = xr
                for i, b in enumerate(blocks):
                    if r == 1 and i == 2:
                        continue
                    h = torch.tanh(b(h))
                # autograd.grad: no .grad accumulation, so the reducer's hooks stay silent
                gs = torch.autograd.grad(h.sum(), params, allow_unused=True)
                for w_, gr in zip(want, gs):
                    if gr is not None:
                        w_ += gr
            ok = all(torch.allclose(g, w_, rtol=1e-5, atol=1e-6) for g, w_ in zip(got, want))
            ret[f"ok_{rank}_{step}"] = bool(ok)
            ret[f"n_{rank}_{step}"] = n
            ret[f"launched_{rank}_{step}"] = launched
            red.launched_in_backward = 0
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_overlapped_reducer_launches_in_order_from_gradient_hooks():
    """overlap=True: buckets are launched from post-accumulate-grad hooks during backward, always in
    bucket order; a rank that skipped a block (no gradient for it) launches that bucket - and the
    ones behind it - from finish(), and the sums still equal the sum of the per-rank gradients."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_overlap_worker, args=(world, port, ret), nprocs=world, join=True)
        r = dict(ret)
        assert all(r[f"ok_{k}_{s}"] for k in range(world) for s in range(2)), r
        assert all(r[f"n_{k}_{s}"] == 4 for k in range(world) for s in range(2)), r
        # rank 0 saw every gradient during backward; rank 1 stalls at the skipped block's bucket
        assert r["launched_0_0"] == 4 and r["launched_1_0"] == 1, r
