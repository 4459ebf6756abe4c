"""
VAD + chunking on the GPU.  Same surface as the reference's VAD classes
(/root/reference/whisperx/vads/vad.py:7-53, pyannote.py:264-301, silero.py:17-69): `preprocess_audio`, `__call__` on
{"waveform", "sample_rate"} and the static `merge_chunks(result, chunk_size, onset, offset)` the pipeline calls
(/root/reference/whisperx/asr.py:97-115) — but the frame scores stay in HBM and Binarize (hysteresis + min-cut) and
Vad.merge_chunks run as one kernel (csrc/wxb_vad.cu); only the final chunk table comes back.

No VAD checkpoint exists offline (Silero is fetched by torch.hub, the pyannote model file is not in the tree): `EnergyVad`
scores frames by log-energy as a stand-in; `GpuVad` takes ANY frame scorer (a callable: device audio -> device scores) with its
frame clock, e.g. a Silero / pyannote segmentation module on the GPU.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np
import torch

from .vad import Vad


@dataclass
class FrameScores:
    """What GpuVad.__call__ returns (the reference's classes return a SlidingWindowFeature / a list of segments)."""
    scores: torch.Tensor          # f32 [n_frames] on the device
    n_samples: int
    frame_duration: float
    frame_step: float
    frame_start: float = 0.0
    ctx: object = None


class GpuVad(Vad):
    def __init__(self, scorer: Callable[[torch.Tensor], torch.Tensor], frame_duration: float, frame_step: float, frame_start: float = 0.0,
                 vad_onset: float = 0.5, vad_offset: Optional[float] = 0.363, chunk_size: float = 30.0, device_index: int = 0, **kwargs):
        super().__init__(vad_onset)
        from .._native import get_context
        self.ctx = get_context(device_index)
        self.scorer = scorer
        self.frame_duration, self.frame_step, self.frame_start = frame_duration, frame_step, frame_start
        self.vad_offset, self.chunk_size = vad_offset, chunk_size

    @staticmethod
    def preprocess_audio(audio):
        return audio

    def __call__(self, audio, **kwargs) -> FrameScores:
        if audio.get("sample_rate", 16000) != 16000:
            raise ValueError("Only 16000Hz sample rate is allowed")  # silero.py:36-38
        wav = audio["waveform"]
        if not (isinstance(wav, torch.Tensor) and wav.is_cuda):
            wav = torch.as_tensor(np.ascontiguousarray(np.asarray(wav, dtype=np.float32).reshape(-1))).to(self.ctx.device)
        wav = wav.reshape(-1).contiguous()
        return FrameScores(self.scorer(wav).contiguous(), wav.numel(), self.frame_duration, self.frame_step, self.frame_start, self.ctx)

    @staticmethod
    def merge_chunks(result: FrameScores, chunk_size, onset: float = 0.5, offset: Optional[float] = None) -> List[dict]:
        """Pyannote.merge_chunks (pyannote.py:282-301) on device-resident scores: [{"start", "end", "segments": [(s, e), ...]}]."""
        assert chunk_size > 0
        (r,) = result.ctx.vad_chunks(result.scores, np.array([0, result.scores.numel()]), np.array([result.n_samples]), chunk_size,
                                     onset=onset, offset=offset, frame_duration=result.frame_duration, frame_step=result.frame_step,
                                     frame_start=result.frame_start)
        if len(r["chunks"]) == 0:
            print("No active speech found in audio")
            return []
        out = []
        for k, (s, e) in enumerate(r["chunks"]):
            members = r["regions"][r["chunk_first"][k]: r["chunk_first"][k + 1]]
            out.append({"start": float(s), "end": float(e), "segments": [(float(a), float(b)) for a, b in members]})
        return out


class EnergyVad(GpuVad):
    """Log-energy frame scores (25 ms frames every 10 ms) computed on the device — a stand-in for a trained VAD model."""

    def __init__(self, floor_db: float = -50.0, width_db: float = 6.0, **kwargs):
        super().__init__(lambda wav: self.ctx.vad_energy_scores(wav, floor_db, width_db), 0.025, 0.010, 0.0, **kwargs)
