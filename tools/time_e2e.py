"""e2e timing variants of EncoderCTCPipeline at the C2 shape (diagnostic)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from oracle import synth  # noqa: E402
from tailored_avsr_b200.ctc.ctc import CTC  # noqa: E402
from tailored_avsr_b200.encoder.branchformer.encoder import MyBranchformerEncoder  # noqa: E402
from tailored_avsr_b200.pipeline import EncoderCTCPipeline  # noqa: E402

dev = torch.device("cuda:0")
w = bench.WORKLOAD if hasattr(bench, "WORKLOAD") else dict(B=32, T=250, feat=512, vocab=41, Lmax=100)
enc = MyBranchformerEncoder(input_size=w["feat"], **bench.enc_cfg())
ctc = CTC(odim=w["vocab"], encoder_output_size=256, dropout_rate=0.0)
synth.fill_module(enc, seed=0)
synth.fill_module(ctc, seed=0, prefix="ctc.")
pipe = EncoderCTCPipeline(enc.to(dev).eval(), ctc.to(dev).eval())
host = [t.pin_memory() for t in bench.make_batch(0)]
for _ in range(3):
    pipe.run(*host)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K = 20


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3


def blocking(fl):
    def f():
        for _ in range(K):
            if fl:
                flush.zero_()
            pipe.run(*host)
    return f


def stream(fl):
    def f():
        for _ in pipe.run_stream(host for _ in range(K)):
            if fl:
                flush.zero_()
    return f


def replay_only():
    for _ in range(K):
        pipe.replay_static()


def h2d_only():
    for _ in range(K):
        for t in host:
            t.to(dev, non_blocking=True)


for name, fn in (("graph replay only", replay_only), ("h2d only", h2d_only),
                 ("blocking, no flush", blocking(False)), ("blocking, flush", blocking(True)),
                 ("stream, no flush", stream(False)), ("stream, flush", stream(True)),
                 ("stream, no flush (again)", stream(False))):
    print(f"{name:28s} {timed(fn):7.3f} ms/step", flush=True)
