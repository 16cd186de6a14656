"""TEST INFRASTRUCTURE (oracle): import stub.  The tailored/conventional AV encoders import
espnet2.asr.ctc.CTC only for a type annotation (src/encoder/audiovisual/tailored/encoder.py:31)."""
import torch


class CTC(torch.nn.Module):
    pass
