"""VAD is upstream of the hot path; only the chunk-merging rule is restated here."""
from .vad import Vad, SegmentX, synthetic_vad_cuts

__all__ = ["Vad", "SegmentX", "synthetic_vad_cuts"]
