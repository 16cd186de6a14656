import sys, traceback, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from tests import _util
from oracle import cases
name = sys.argv[1] if len(sys.argv) > 1 else "asr_small"
enc, ctc, sd = _util.build_dropin(name)
inp = cases.make_inputs(name)
enc = enc.cuda()
try:
    with torch.no_grad():
        y, olens, _ = enc(inp["x"].cuda(), inp["lens"].cuda())
    torch.cuda.synchronize()
    print("OK", y.shape, float(y.abs().max()))
except Exception:
    traceback.print_exc(limit=6)
