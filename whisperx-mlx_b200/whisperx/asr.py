"""
ASR entry points of the B200 build: `load_model` with the reference's signature
(whisperx/asr.py:150-166) and a pipeline object with the reference's `transcribe`
(whisperx/asr.py:28-120: VAD -> cut -> backend.transcribe_batch).
"""
from typing import Dict, List, Optional, Union

import numpy as np

from .audio import SAMPLE_RATE, load_audio
from .types import TranscriptionResult
from .vads import Vad, synthetic_vad_cuts


class B200WhisperPipeline:
    """Same surface as the reference's MLXWhisperPipeline (asr.py:19-147)."""

    def __init__(self, backend, vad_model=None, vad_options: Optional[dict] = None):
        self.backend = backend
        self.vad_model = vad_model
        self.vad_options = vad_options or {}

    def transcribe(self, audio: Union[str, np.ndarray], batch_size: int = 8, chunk_size: int = 30,
                   print_progress: bool = False, combined_progress: bool = False, verbose: bool = False,
                   **kwargs) -> TranscriptionResult:
        if isinstance(audio, str):
            audio = load_audio(audio)
        if hasattr(self.backend, "_align_words") and kwargs.get("word_timestamps", False):
            kwargs["align_words"] = True
        kwargs.pop("word_timestamps", None)
        if self.vad_model is None:
            return self.backend.transcribe(audio, batch_size=batch_size, num_workers=0, chunk_size=chunk_size,
                                           print_progress=print_progress, combined_progress=combined_progress,
                                           verbose=verbose, **kwargs)
        segments = self._segment_audio_with_vad(audio, chunk_size)
        for seg in segments:  # the reference adds the "audio" key in place too (asr.py:70-73)
            seg["audio"] = audio[int(seg["start"] * SAMPLE_RATE): int(seg["end"] * SAMPLE_RATE)]
        return self.backend.transcribe_batch(segments, batch_size=batch_size, print_progress=print_progress,
                                             combined_progress=combined_progress, verbose=verbose, **kwargs)

    def _segment_audio_with_vad(self, audio: np.ndarray, chunk_size: int) -> List[Dict]:
        if self.vad_model == "uniform":
            return synthetic_vad_cuts(len(audio) / SAMPLE_RATE, "uniform", chunk_size=chunk_size)
        vad = self.vad_model
        waveform = vad.preprocess_audio(audio) if hasattr(vad, "preprocess_audio") else audio
        speech = vad({"waveform": waveform, "sample_rate": SAMPLE_RATE})
        merge = vad.merge_chunks if hasattr(vad, "merge_chunks") else Vad.merge_chunks
        return merge(speech, chunk_size, onset=getattr(vad, "vad_onset", 0.5), offset=getattr(vad, "vad_offset", 0.363))

    def detect_language(self, audio: np.ndarray) -> str:
        return self.backend.detect_language(audio)


def load_model(whisper_arch: str, device: str = "cuda", device_index: int = 0, compute_type: str = "bfloat16",
               asr_options: Optional[dict] = None, language: Optional[str] = None, vad_method: Optional[str] = "uniform",
               vad_options: Optional[dict] = None, task: str = "transcribe", download_root: Optional[str] = None,
               local_files_only: bool = False, threads: int = 4, backend: str = "auto", batch_size: int = 8,
               vad_model=None, **kwargs) -> B200WhisperPipeline:
    """whisperx.load_model(name, backend="b200").  `backend` accepts "auto" | "b200" | "cuda_b200";
    the reference's MLX backend names are rejected (this package ships the CUDA backend only).

    `vad_method`: "uniform" (fixed chunk_size cuts, the benchmark's synthetic VAD), "energy" (log-energy frame scores +
    Binarize + merge_chunks on the GPU, vads/gpu.py), None / "none" (no cutting), or pass a ready `vad_model` object
    (callable returning speech regions or device frame scores, with the reference's merge_chunks; vads.GpuVad wraps any
    frame scorer).  The reference's Silero / pyannote checkpoints need network access and are not bundled."""
    if backend not in ("auto", "b200", "cuda_b200"):
        raise ValueError(f"backend '{backend}' is not available in the B200 build (use backend='b200')")
    from .backends.b200 import B200WhisperBackend
    be = B200WhisperBackend(whisper_arch, device=device, device_index=device_index, compute_type=compute_type,
                            download_root=download_root, local_files_only=local_files_only, threads=threads,
                            asr_options=asr_options, language=language, task=task, **kwargs)
    opts = {"chunk_size": 30, "vad_onset": 0.500, "vad_offset": 0.363}
    if vad_options:
        opts.update(vad_options)
    if vad_model is None:
        if vad_method in (None, "none"):
            vad_model = None
        elif vad_method == "uniform":
            vad_model = "uniform"
        elif vad_method == "energy":
            # frame scores, Binarize (hysteresis + min-cut) and merge_chunks on the GPU (vads/gpu.py); the scorer is a log-energy
            # stand-in for the reference's Silero / pyannote models, whose checkpoints need network access
            from .vads.gpu import EnergyVad
            vad_model = EnergyVad(vad_onset=opts["vad_onset"], vad_offset=opts["vad_offset"], chunk_size=opts["chunk_size"],
                                  device_index=device_index)
        else:
            raise RuntimeError(f"vad_method='{vad_method}' needs a VAD model that is outside this build's hot-path scope; "
                               "pass vad_model=<object> or use vad_method='uniform' / None")
    return B200WhisperPipeline(be, vad_model, opts)
