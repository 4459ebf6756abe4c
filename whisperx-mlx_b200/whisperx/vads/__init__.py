"""VAD post-processing and chunking: the chunk-merging rule on the host (vad.py) and Binarize + merge on the
GPU for device-resident frame scores (gpu.py, imported lazily: it needs the CUDA library)."""
from .vad import Vad, SegmentX, synthetic_vad_cuts

__all__ = ["Vad", "SegmentX", "synthetic_vad_cuts", "GpuVad", "EnergyVad", "FrameScores"]


def __getattr__(name):
    if name in ("GpuVad", "EnergyVad", "FrameScores"):
        from . import gpu
        return getattr(gpu, name)
    raise AttributeError(name)
