"""TEST INFRASTRUCTURE (oracle): espnet2 AbsEncoder restated (abstract base of all encoders)."""
from abc import ABC, abstractmethod

import torch


class AbsEncoder(torch.nn.Module, ABC):
    @abstractmethod
    def output_size(self) -> int:
        raise NotImplementedError

    @abstractmethod
    def forward(self, xs_pad, ilens, prev_states=None):
        raise NotImplementedError
