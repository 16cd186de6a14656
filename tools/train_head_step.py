"""Training-mode step of what is trainable on the B200 path today (SURVEY.md §8d C5 "training step",
§8e): encoder forward (frozen: its backward kernels are not built yet) -> CTC head + loss forward
-> CUDA backward of loss and head (d ctc_lo.weight / bias, d hs) -> bucketed NCCL all-reduce of the
gradients (parallel.GradBucketReducer).  Utterances are sharded over the ranks, the loss is
normalised by the GLOBAL batch, and rank 0 checks the reduced gradients against the full-batch
gradients it computes alone.  Prints one JSON line.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_head_step.py
"""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tailored_avsr_b200 import ops, parallel  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warm = 20, 5
    bench.select_workload(types.SimpleNamespace(workload="C2", batch=0, T=0))
    w = bench.WORKLOAD
    enc, _, ctc, _ = bench.build_modules()
    enc, ctc = enc.to(dev).eval(), ctc.to(dev).train()
    Bg = w["B"] * world                                   # global batch, identical on every rank
    from oracle import synth
    feats = synth.randn((Bg, w["T"], w["feat"]), 3)
    lens = torch.full((Bg,), w["T"], dtype=torch.int64)
    ys = synth.rand_targets(Bg, w["Lmax"], w["vocab"], 4)
    ylens = torch.full((Bg,), w["Lmax"], dtype=torch.int64)
    mine = torch.tensor(parallel.shard_utterances(lens.tolist(), world, rank))
    f, l, y, yl = (t[mine].to(dev) for t in (feats, lens, ys, ylens))
    reducer = parallel.GradBucketReducer(list(ctc.parameters()), bucket_mb=25.0)

    def step():
        ctc.zero_grad(set_to_none=True)
        with torch.no_grad():
            hs, olens, _ = enc(f, l)
        hs = hs.detach().requires_grad_(True)
        ctc.reduce = False
        vec = ctc(hs, olens, y, yl) * len(mine)           # nll_b of the local utterances
        ctc.reduce = True
        loss = vec.sum() / Bg                              # ctc.py:62-66 with the GLOBAL batch
        loss.backward()
        n = reducer.reduce()
        return parallel.global_ctc_loss(vec.detach(), Bg), hs.grad, n

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, dhs, ncoll = step()
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    launches = (ops.launch_count() - l0) // steps
    ok = None
    if rank == 0:
        # full-batch reference gradients on one GPU through the same CUDA path
        got_w, got_b = ctc.ctc_lo.weight.grad.clone(), ctc.ctc_lo.bias.grad.clone()
        ctc.zero_grad(set_to_none=True)
        with torch.no_grad():
            hs_all, ol_all, _ = enc(feats.to(dev), lens.to(dev))
        full = ctc(hs_all, ol_all, ys.to(dev), ylens.to(dev))
        full.backward()
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())  # noqa: E731
        ok = {"loss": abs(float(loss) - float(full)) <= 1e-5 * abs(float(full)),
              "dW": rel(got_w, ctc.ctc_lo.weight.grad), "db": rel(got_b, ctc.ctc_lo.bias.grad)}
        print(json.dumps({"what": "C2 training-mode step: frozen encoder fwd + CTC head/loss fwd+bwd "
                                  "(CUDA) + bucketed gradient all-reduce",
                          "n_gpus": world, "global_batch": Bg, "ms_per_step": round(ms, 4),
                          "frames_per_s": round(Bg * w["T"] / (ms * 1e-3)),
                          "kernels_per_step": launches, "collectives_per_step": ncoll,
                          "check_vs_full_batch": ok, "loss": float(loss)}), flush=True)
        assert ok["loss"] and ok["dW"] < 1e-4 and ok["db"] < 1e-4, ok
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
