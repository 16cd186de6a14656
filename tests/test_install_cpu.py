"""The install hook (tailored_avsr_b200/install.py) executed on a stand-in for the reference's
task modules: espnet2 `ClassChoices` registries are objects with a `.classes` dict keyed by the
YAML strings (src/tasks/asr.py:145-166, src/tasks/avsr.py:140-172) and `CTC` is a module-level
name (asr.py:12, avsr.py:43)."""
import pytest

from tailored_avsr_b200 import install
from tailored_avsr_b200.audiovisual_fusion.adaptive_audiovisual_fusion import AdaptiveAudioVisualFusion
from tailored_avsr_b200.ctc.ctc import CTC
from tailored_avsr_b200.embedding_for_avsr.default import DefaultEmbeddingLayerForAVSR
from tailored_avsr_b200.encoder.audiovisual.conventional.encoder import ConventionalEncoder
from tailored_avsr_b200.encoder.audiovisual.tailored.encoder import TailoredEncoder
from tailored_avsr_b200.encoder.branchformer.encoder import MyBranchformerEncoder


class FakeClassChoices:
    """The part of espnet2.train.class_choices.ClassChoices the hook touches."""

    def __init__(self, name, classes, type_check=None):
        self.name = name
        self.classes = dict(classes)
        self.type_check = type_check

    def get_class(self, name):
        return self.classes[name.lower()]


class _RefEncoder:  # what the registries hold before the hook runs
    pass


def test_install_asr_rebinds_encoder_registry_and_ctc():
    ns = {"encoder_choices": FakeClassChoices("encoder", {"branchformer": _RefEncoder, "other": _RefEncoder}),
          "CTC": object}
    install.install_asr(ns)
    assert ns["encoder_choices"].get_class("branchformer") is MyBranchformerEncoder
    assert ns["encoder_choices"].get_class("other") is _RefEncoder          # untouched
    assert ns["CTC"] is CTC
    # the task instantiates `encoder_class(input_size=input_size, **args.encoder_conf)` (asr.py:545)
    enc = ns["encoder_choices"].get_class("branchformer")(input_size=80, num_blocks=1, input_layer="conv2d",
                                                          ffn_activation_type="swish")
    assert enc.output_size() == 256
    ctc = ns["CTC"](odim=41, encoder_output_size=enc.output_size(), dropout_rate=0.0)
    assert ctc.ctc_lo.weight.shape == (41, 256)


def test_install_avsr_rebinds_every_registry():
    ns = {"encoder_choices": FakeClassChoices("encoder", {"tailored": _RefEncoder, "conventional": _RefEncoder}),
          "audiovisual_fusion_choices": FakeClassChoices("audiovisual_fusion", {"adaptive": _RefEncoder}),
          "acoustic_embed_choices": FakeClassChoices("acoustic_embed", {"default": _RefEncoder}),
          "visual_embed_choices": FakeClassChoices("visual_embed", {"default": _RefEncoder}),
          "CTC": object}
    install.install_avsr(ns)
    assert ns["encoder_choices"].get_class("tailored") is TailoredEncoder
    assert ns["encoder_choices"].get_class("conventional") is ConventionalEncoder
    assert ns["audiovisual_fusion_choices"].get_class("adaptive") is AdaptiveAudioVisualFusion
    assert ns["acoustic_embed_choices"].get_class("default") is DefaultEmbeddingLayerForAVSR
    assert ns["visual_embed_choices"].get_class("default") is DefaultEmbeddingLayerForAVSR
    assert ns["CTC"] is CTC
    # avsr.py:618-631: encoder_class(embed_pos_enc_layer_type=..., embed_rel_pos_type=..., **conf)
    enc = ns["encoder_choices"].get_class("tailored")(
        embed_pos_enc_layer_type="rel_pos", embed_rel_pos_type="latest", num_blocks=1,
        acoustic_use_attn=[True], visual_use_attn=[False])
    assert enc.output_size() == 256


def test_install_rejects_a_namespace_without_registries():
    with pytest.raises(KeyError):
        install.install_asr({})
    with pytest.raises(TypeError):
        install.install_asr({"encoder_choices": object(), "CTC": object})


def test_install_only_selects_a_subset():
    """A training run can keep the stock classes where the B200 path is inference-only."""
    ns = {"encoder_choices": FakeClassChoices("encoder", {"tailored": _RefEncoder, "conventional": _RefEncoder}),
          "audiovisual_fusion_choices": FakeClassChoices("audiovisual_fusion", {"adaptive": _RefEncoder}),
          "acoustic_embed_choices": FakeClassChoices("acoustic_embed", {"default": _RefEncoder}),
          "visual_embed_choices": FakeClassChoices("visual_embed", {"default": _RefEncoder}),
          "CTC": object}
    install.install_avsr(ns, only=("conventional", "ctc"))
    assert ns["encoder_choices"].get_class("conventional") is ConventionalEncoder
    assert ns["encoder_choices"].get_class("tailored") is _RefEncoder
    assert ns["audiovisual_fusion_choices"].get_class("adaptive") is _RefEncoder
    assert ns["acoustic_embed_choices"].get_class("default") is _RefEncoder
    assert ns["CTC"] is CTC
    with pytest.raises(ValueError):
        install.install_asr({"encoder_choices": FakeClassChoices("encoder", {}), "CTC": object}, only=("nope",))
