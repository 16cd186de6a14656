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
