"""Install hook: rebind the reference's registry entries to the B200 drop-in classes.

The reference selects its encoder through espnet2 `ClassChoices` registries and binds `CTC` at
import time (src/tasks/asr.py:12,145-166,589-591; src/tasks/avsr.py:43,156-164,683).  A maintainer
adds to each task file

    from tailored_avsr_b200.install import install_asr      # or install_avsr
    install_asr(globals())

and `avsr_main.py`, the YAML configs and checkpoints run unchanged.

Training coverage of what gets installed (INTEGRATION.md "Training"): grad-mode calls of every
class are built (training.py), InterCTC taps and self-conditioning included.  `only=` installs a
subset, should a
maintainer want a stock class somewhere: `install_avsr(globals(), only=("conventional", "ctc"))`.
"""
from __future__ import annotations

from .audiovisual_fusion.adaptive_audiovisual_fusion import AdaptiveAudioVisualFusion
from .ctc.ctc import CTC
from .embedding_for_avsr.default import DefaultEmbeddingLayerForAVSR
from .encoder.audiovisual.conventional.encoder import ConventionalEncoder
from .encoder.audiovisual.tailored.encoder import TailoredEncoder
from .encoder.branchformer.encoder import MyBranchformerEncoder


def _classes(registry):
    classes = getattr(registry, "classes", None)
    if not isinstance(classes, dict):
        raise TypeError("expected an espnet2 ClassChoices registry with a `.classes` dict")
    return classes


ASR_PARTS = ("branchformer", "ctc")
AVSR_PARTS = ("tailored", "conventional", "fusion", "embed", "ctc")


def _select(only, known):
    if only is None:
        return set(known)
    unknown = set(only) - set(known)
    if unknown:
        raise ValueError(f"install: unknown part(s) {sorted(unknown)}; choose from {known}")
    return set(only)


def install_asr(namespace: dict, only=None) -> None:
    """`namespace` is the globals() of src/tasks/asr.py; `only`: subset of ASR_PARTS."""
    parts = _select(only, ASR_PARTS)
    if "branchformer" in parts:
        _classes(namespace["encoder_choices"])["branchformer"] = MyBranchformerEncoder
    if "ctc" in parts:
        namespace["CTC"] = CTC


def install_avsr(namespace: dict, only=None) -> None:
    """`namespace` is the globals() of src/tasks/avsr.py; `only`: subset of AVSR_PARTS."""
    parts = _select(only, AVSR_PARTS)
    if parts & {"tailored", "conventional"}:
        classes = _classes(namespace["encoder_choices"])
        if "tailored" in parts:
            classes["tailored"] = TailoredEncoder
        if "conventional" in parts:
            classes["conventional"] = ConventionalEncoder
    if "fusion" in parts:
        # src/tasks/avsr.py:165-172: audiovisual_fusion_choices, key "adaptive"
        _classes(namespace["audiovisual_fusion_choices"])["adaptive"] = AdaptiveAudioVisualFusion
    if "embed" in parts:
        # src/tasks/avsr.py:140-155: acoustic_embed_choices / visual_embed_choices, key "default"
        _classes(namespace["acoustic_embed_choices"])["default"] = DefaultEmbeddingLayerForAVSR
        _classes(namespace["visual_embed_choices"])["default"] = DefaultEmbeddingLayerForAVSR
    if "ctc" in parts:
        namespace["CTC"] = CTC
