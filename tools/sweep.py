"""Throughput sweep (inference part): encoder + CTC valid frames/s over
T x batch for one workload, device-resident CUDA-graph replay, CUDA-event timed with an L2 flush
between steps (same timing rules as bench.py, whose builders it reuses).  One JSON line per point.

  python tools/sweep.py --workload C2 --T 250,500,1000,1500 --batch 1,8,32,128,256 [--steps 10]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py ...
Under torchrun every rank runs its own batch of each point (utterance sharding, `batch` is per GPU),
the step time is the max over ranks and frames/s the whole-job aggregate; rank 0 prints.
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--T", default="250,500,1000,1500")
    ap.add_argument("--batch", default="1,8,32,128,256")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--max-rows", type=int, default=400000, help="skip points with B*T above this")
    ap.add_argument("--dtype", default="bf16", choices=["tf32", "tf32x3", "bf16"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: the CUDA-graph training step of bench.py (forward + loss + backward "
                         "+ gradient all-reduce, every dropout active) instead of the inference step")
    a = ap.parse_args()
    from tailored_avsr_b200 import engine
    from tailored_avsr_b200.pipeline import AVEncoderCTCPipeline, EncoderCTCPipeline
    engine.set_compute_dtype(a.dtype)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    mods = None
    for T in [int(x) for x in a.T.split(",")]:
        for B in [int(x) for x in a.batch.split(",")]:
            if B * T > a.max_rows:
                continue
            bench.select_workload(types.SimpleNamespace(workload=a.workload, batch=B, T=T))
            if mods is None:  # weights do not depend on (B, T)
                enc, fusion, ctc, _ = bench.build_modules()
                mods = (enc.to(dev).eval(), fusion.to(dev).eval() if fusion is not None else None,
                        ctc.to(dev).eval())
            enc, fusion, ctc = mods
            if a.mode == "train":
                targs = types.SimpleNamespace()
                res = bench.train_graph_leg(targs, enc, ctc, rank, world, dev, dist, steps=a.steps,
                                            warmup=max(1, a.warmup), fusion=fusion)
                if rank == 0:
                    res = res or {}
                    print(json.dumps({"workload": a.workload, "mode": "train", "B": B, "T": T,
                                      "ms_per_step": res.get("ms_per_step"), "frames_per_s": res.get("value"),
                                      "unavailable": res.get("unavailable"), "n_gpus": world,
                                      "batch_per_gpu": B, "step": res.get("step")}), flush=True)
                torch.cuda.empty_cache()
                continue
            pipe = (EncoderCTCPipeline(enc, ctc) if fusion is None
                    else AVEncoderCTCPipeline(enc, fusion, ctc))
            host, frames = bench.make_batch(rank)
            batch_dev = [t.to(dev) for t in host]
            for _ in range(a.warmup):
                res = pipe.run_device(*batch_dev)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            evs = []
            for _ in range(a.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pipe.replay_static()
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = sum(x.elapsed_time(y) for x, y in evs) / a.steps
            tot = torch.tensor([float(frames)], device=dev, dtype=torch.float64)
            if dist is not None:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.all_reduce(tot)
                ms = float(t[0])
            fps = float(tot[0]) / (ms * 1e-3)
            if rank == 0:
                print(json.dumps({"workload": a.workload, "B": B, "T": T, "valid_frames": int(tot[0]),
                                  "ms_per_step": round(ms, 4), "frames_per_s": round(fps),
                                  "model_tflops": round(fps * bench.flops_per_frame() / 1e12, 1),
                                  "loss": float(res["loss"]), "dtype": a.dtype, "n_gpus": world,
                                  "batch_per_gpu": B}), flush=True)
            del pipe, batch_dev, res
            torch.cuda.empty_cache()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
