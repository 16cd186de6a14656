"""One eager training step at the C2 shape between cudaProfilerStart / Stop, for an ncu launch list:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/train_launches.csv python tools/prof_train.py [--eval]
  python tools/ncu_launches.py gpurun_out/train_launches.csv

Without ncu it prints the step's CUDA-event time split into forward / loss / backward."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tailored_avsr_b200 import engine  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    bench.select_workload(types.SimpleNamespace(workload="C2", batch=0, T=0))
    w = bench.WORKLOAD
    enc, _, ctc, _ = bench.build_modules()
    enc, ctc = enc.to(dev), ctc.to(dev)
    enc.train("--eval" not in sys.argv)
    engine.set_compute_dtype("tf32")
    host, frames = bench.make_batch(0)
    feats, lens, ys, ylens = (t.to(dev) for t in host)
    params = list(enc.parameters()) + list(ctc.parameters())
    for p in params:
        p.requires_grad_(True)

    def step(timed=False):
        for p in params:
            p.grad = None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        out, olens, _ = enc(feats, lens)
        ev[1].record()
        loss = ctc(out, olens, ys, ylens)
        ev[2].record()
        loss.backward()
        ev[3].record()
        torch.cuda.synchronize()
        if timed:
            print("forward %.2f ms, ctc %.2f ms, backward %.2f ms, loss %.4f" % (
                ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), float(loss)))

    for _ in range(3):
        step()
    step(timed=True)
    if "--cprofile" in sys.argv:       # where the HOST time of a step goes (forward thread only)
        import cProfile
        import pstats
        import time
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        step()
        pr.disable()
        print("host wall of the profiled step: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    torch.cuda.profiler.start()
    step()
    torch.cuda.profiler.stop()
    step(timed=True)


if __name__ == "__main__":
    main()
