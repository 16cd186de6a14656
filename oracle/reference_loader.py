"""TEST INFRASTRUCTURE: import the REAL reference modules (src/encoder/**, src/ctc/ctc.py under
/root/reference) on top of oracle/espnet_shim.  Works only where /root/reference exists (the build
container); the GPU box uses the golden vectors generated from it (oracle/gen_golden.py)."""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("TAVSR_REFERENCE_ROOT", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "espnet_shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "encoder"))


def load():
    """Returns a namespace with the reference classes.  typeguard 4.x rejects the reference's
    `ignore_nan_grad: bool = None` default (src/ctc/ctc.py:28), so check_argument_types is
    neutralised first (SURVEY.md probe table)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import typeguard
    typeguard.check_argument_types = lambda *a, **k: True
    for p in (REFERENCE_ROOT, SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    from types import SimpleNamespace

    from src.audiovisual_fusion.adaptive_audiovisual_fusion import AdaptiveAudioVisualFusion
    from src.ctc.ctc import CTC
    from src.ctc.interctc_residual_module import InterCTCResidualModule
    from src.encoder.audiovisual.conventional.encoder import ConventionalEncoder
    from src.encoder.audiovisual.tailored.encoder import TailoredEncoder
    from src.encoder.audiovisual.tailored.encoder_layer import TailoredEncoderLayer
    from src.encoder.branchformer.encoder import MyBranchformerEncoder
    from src.encoder.branchformer.encoder_layer import MyBranchformerEncoderLayer
    return SimpleNamespace(CTC=CTC, AdaptiveAudioVisualFusion=AdaptiveAudioVisualFusion, InterCTCResidualModule=InterCTCResidualModule,
                           ConventionalEncoder=ConventionalEncoder,
                           TailoredEncoder=TailoredEncoder, TailoredEncoderLayer=TailoredEncoderLayer,
                           MyBranchformerEncoder=MyBranchformerEncoder,
                           MyBranchformerEncoderLayer=MyBranchformerEncoderLayer)
