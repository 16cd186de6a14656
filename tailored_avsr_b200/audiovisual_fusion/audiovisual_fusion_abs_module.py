"""Abstract base of the audio-visual fusion modules, same contract as the reference's
src/audiovisual_fusion/audiovisual_fusion_abs_module.py:6-19 (the espnet2 `ClassChoices` registry
of src/tasks/avsr.py:165-172 type-checks against it)."""
from abc import ABC, abstractmethod
from typing import Tuple

import torch


class AudioVisualFusionAbsModule(torch.nn.Module, ABC):
    @abstractmethod
    def output_size(self) -> int:
        raise NotImplementedError

    @abstractmethod
    def forward(self, audio_pad: torch.Tensor, audio_masks: torch.Tensor, video_pad: torch.Tensor,
                video_masks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError
