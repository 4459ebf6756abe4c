"""
Chunk merging of VAD speech regions (reference: whisperx/vads/vad.py:20-53 `Vad.merge_chunks`) and
the deterministic synthetic VAD cuts the benchmark uses in place of a VAD model.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class SegmentX:
    start: float
    end: float
    speaker: Optional[str] = None


class Vad:
    def __init__(self, vad_onset: float):
        if not (0 < vad_onset < 1):
            raise ValueError("vad_onset is a decimal value between 0 and 1.")
        self.vad_onset = vad_onset

    @staticmethod
    def preprocess_audio(audio):
        return audio

    @staticmethod
    def merge_chunks(segments, chunk_size, onset: float = 0.5, offset: Optional[float] = None) -> List[dict]:
        """Greedy left-to-right merge: a new chunk starts when adding the next speech region would
        make the current chunk longer than `chunk_size` seconds (and the chunk is non-empty)."""
        merged = []
        if not segments:
            return merged
        chunk_start = segments[0].start
        chunk_end = 0
        members = []
        for seg in segments:
            too_long = seg.end - chunk_start > chunk_size
            if too_long and chunk_end - chunk_start > 0:
                merged.append({"start": chunk_start, "end": chunk_end, "segments": members})
                chunk_start, members = seg.start, []
            chunk_end = seg.end
            members.append((seg.start, seg.end))
        merged.append({"start": chunk_start, "end": chunk_end, "segments": members})
        return merged


def synthetic_vad_cuts(total_seconds: float, mode: str = "uniform", seed: int = 1234, chunk_size: float = 30.0):
    """`uniform`: back-to-back chunk_size cuts; `ragged`: durations ~ U(5, chunk_size) (seeded)."""
    cuts, t = [], 0.0
    rng = np.random.RandomState(seed)
    while t < total_seconds - 1e-9:
        dur = chunk_size if mode == "uniform" else float(rng.uniform(5.0, chunk_size))
        end = min(t + dur, total_seconds)
        cuts.append({"start": round(t, 3), "end": round(end, 3), "segments": [(round(t, 3), round(end, 3))]})
        t = end
    return cuts
