"""TEST INFRASTRUCTURE (oracle): import stub; the conformer alternative is out of scope."""
from espnet2.asr.encoder.abs_encoder import AbsEncoder


class ConformerEncoder(AbsEncoder):
    def __init__(self, *a, **k):
        raise NotImplementedError("ConformerEncoder is outside the restated surface")

    def output_size(self):
        raise NotImplementedError

    def forward(self, xs_pad, ilens, prev_states=None):
        raise NotImplementedError
