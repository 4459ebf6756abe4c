"""
ORACLE — test infrastructure only.  Never imported by the product path (whisperx-mlx_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

CPU restatement of the reference's VAD post-processing, pinned bit-exactly to outputs of the reference's own
code (tests/golden/vad_golden.npz, tests/golden/make_vad_golden.py):

    /root/reference/whisperx/vads/pyannote.py:134-216   Binarize.__call__: hysteresis thresholding of frame scores with the
                                                         WhisperX min-cut (a region longer than max_duration is divided at the
                                                         lowest score of its second half)
    /root/reference/whisperx/vads/pyannote.py:282-301   Pyannote.merge_chunks = Binarize(max_duration=chunk_size) -> timeline
    /root/reference/whisperx/vads/vad.py:20-53          Vad.merge_chunks: greedy merge of speech regions into <= chunk_size chunks

The frame clock follows pyannote.core (un-vendored, uv.lock pyannote-core 5.0.0): frame i spans [start + i step, + duration),
its timestamp is the middle 0.5 (s + e); a region shorter than 1e-6 s is empty and never stored.

The reference keeps `curr_scores` / `curr_timestamps` lists whose first element can be STALE (the frame at which the previous
region closed, or frame 0); `members` below holds the frame indices behind those lists.

`energy_scores` is NOT a restatement of the reference: no VAD checkpoint exists offline (Silero comes from torch.hub, the
pyannote model file is not in the tree), so a log-energy frame scorer stands in as the source of frame scores on the GPU path.
"""
from typing import List, Optional, Tuple

import numpy as np


def frame_times(n: int, duration: float, step: float, start: float = 0.0) -> List[float]:
    out = []
    for i in range(n):
        s = start + i * step
        out.append(0.5 * (s + (s + duration)))
    return out


def binarize(scores: np.ndarray, duration: float, step: float, start: float, onset: float = 0.5, offset: Optional[float] = None,
             max_duration: float = float("inf")) -> List[Tuple[float, float]]:
    """Speech regions (start, end) of one score track, in the order the reference's timeline lists them."""
    y = np.asarray(scores, dtype=np.float32)
    offset = offset or onset
    t = frame_times(len(y), duration, step, start)
    regions = []
    region_start = t[0]
    active = bool(y[0] > onset)
    members = [0]  # frames behind the reference's curr_scores / curr_timestamps lists
    for i in range(1, len(y)):
        if active:
            if t[i] - region_start > max_duration:
                half = len(members) // 2
                cut = half + int(np.argmin(y[members[half:]]))  # first minimum of the second half
                regions.append((region_start, t[members[cut]]))
                region_start = t[members[cut]]
                members = members[cut + 1:]
            elif y[i] < offset:
                regions.append((region_start, t[i]))
                region_start = t[i]
                active = False
                members = []
            members.append(i)  # in every sub-case: after a closing frame the list is [that frame] and goes stale
        elif y[i] > onset:
            region_start = t[i]
            active = True      # the lists are not reset: the stale element stays in front of the frames that follow
    if active:
        regions.append((region_start, t[-1]))
    regions = [r for r in regions if (r[1] - r[0]) > 1e-6]
    return sorted(set(regions))


def merge_chunks(segments: List[Tuple[float, float]], chunk_size: float) -> List[dict]:
    """vad.py:20-53 on (start, end) pairs."""
    if not segments:
        return []
    chunks = []
    cur_start, cur_end, members = segments[0][0], 0, []
    for (s, e) in segments:
        if e - cur_start > chunk_size and cur_end - cur_start > 0:
            chunks.append({"start": cur_start, "end": cur_end, "segments": members})
            cur_start, members = s, []
        cur_end = e
        members.append((s, e))
    chunks.append({"start": cur_start, "end": cur_end, "segments": members})
    return chunks


def vad_chunks(scores, duration, step, start, chunk_size, onset=0.5, offset=None):
    """Pyannote.merge_chunks (pyannote.py:282-301)."""
    return merge_chunks(binarize(scores, duration, step, start, onset, offset, max_duration=chunk_size), chunk_size)


# ---- stand-in frame scorer (not in the reference) -------------------------------------------------------------------------
ENERGY_WIN, ENERGY_HOP = 400, 160  # 25 ms / 10 ms at 16 kHz


def energy_scores(audio: np.ndarray, floor_db: float = -50.0, width_db: float = 6.0) -> np.ndarray:
    """score[i] = sigmoid((10 log10(mean(x^2 over frame i) + 1e-10) - floor_db) / width_db), frames of 400 samples every 160
    (zero-padded at the end), float32 arithmetic except the mean (float64 accumulation)."""
    x = np.asarray(audio, dtype=np.float32)
    n = max(1, (len(x) + ENERGY_HOP - 1) // ENERGY_HOP)
    xp = np.zeros((n - 1) * ENERGY_HOP + ENERGY_WIN, dtype=np.float32)
    xp[:len(x)] = x
    frames = np.lib.stride_tricks.sliding_window_view(xp, ENERGY_WIN)[::ENERGY_HOP][:n]
    e = (frames.astype(np.float64) ** 2).mean(axis=1)
    db = (10.0 * np.log10(e + 1e-10)).astype(np.float32)
    return (1.0 / (1.0 + np.exp(-(db - np.float32(floor_db)) / np.float32(width_db)))).astype(np.float32)
