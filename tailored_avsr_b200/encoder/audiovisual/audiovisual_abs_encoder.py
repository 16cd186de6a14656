"""Abstract base of the audio-visual encoders (reference:
src/encoder/audiovisual/audiovisual_abs_encoder.py:6-20).  When the reference package is importable
its own ABC is used, so `ClassChoices(type_check=AudioVisualAbsEncoder)` (src/tasks/avsr.py:162)
accepts the drop-in."""
from abc import ABC, abstractmethod
from typing import Optional, Tuple

import torch

try:  # pragma: no cover - only when running inside the reference tree
    from src.encoder.audiovisual.audiovisual_abs_encoder import AudioVisualAbsEncoder  # type: ignore
except Exception:  # noqa: BLE001
    class AudioVisualAbsEncoder(torch.nn.Module, ABC):
        @abstractmethod
        def output_size(self) -> int:
            raise NotImplementedError

        @abstractmethod
        def forward(self, audio_pad, audio_ilens, video_pad, video_ilens, prev_states=None
                    ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
            raise NotImplementedError
