"""
ORACLE — test infrastructure only.  Never imported by the product path (whisperx-mlx_b200/);
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

CPU restatement (numpy, fp32 arithmetic made explicit) of the reference CTC forced aligner:
    /root/reference/whisperx/alignment.py:387-404  get_trellis
    /root/reference/whisperx/alignment.py:407-437  get_wildcard_emission
    /root/reference/whisperx/alignment.py:447-481  backtrack
    /root/reference/whisperx/alignment.py:500-579  backtrack_beam
    /root/reference/whisperx/alignment.py:597-613  merge_repeats

Pinned: tests/golden/ctc_*.npz hold trellises and paths produced by the reference functions
themselves (tests/golden/make_golden.py); tests/test_oracle_cpu.py checks bit-equality.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

F32 = np.float32
NEG_INF = F32(-np.inf)
POS_INF = F32(np.inf)


@dataclass
class Point:  # alignment.py:440-444
    token_index: int
    time_index: int
    score: float


@dataclass
class Segment:  # alignment.py:583-595
    label: str
    start: int
    end: int
    score: float

    @property
    def length(self):
        return self.end - self.start


def wildcard_emission(frame: np.ndarray, tokens: np.ndarray, blank_id: int) -> np.ndarray:
    """alignment.py:407-437: token -1 scores as the best non-blank label of the frame."""
    tokens = np.asarray(tokens, dtype=np.int64)
    regular = frame[np.clip(tokens, 0, None)]
    masked = frame.copy()
    masked[blank_id] = NEG_INF
    return np.where(tokens == -1, masked.max(), regular).astype(F32)


def get_trellis(emission: np.ndarray, tokens, blank_id: int = 0) -> np.ndarray:
    """alignment.py:387-404.  emission f32 [T,V] log-probs; returns f32 [T,N]."""
    emission = np.asarray(emission, dtype=F32)
    tokens = np.asarray(tokens, dtype=np.int64)
    T, N = emission.shape[0], len(tokens)
    trellis = np.zeros((T, N), dtype=F32)
    # torch CPU cumsum on fp32 accumulates in fp64 and rounds each element (alignment.py:392)
    trellis[1:, 0] = np.cumsum(emission[1:, blank_id].astype(np.float64)).astype(F32)
    trellis[0, 1:] = NEG_INF
    # alignment.py:394 `trellis[-num_tokens + 1:, 0] = inf` (Python slice semantics; N == 1 -> [0:])
    trellis[slice(-N + 1, None) if N != 1 else slice(0, None), 0] = POS_INF
    for t in range(T - 1):
        stay = trellis[t, 1:] + emission[t, blank_id]
        change = trellis[t, :-1] + wildcard_emission(emission[t], tokens[1:], blank_id)
        # torch.maximum propagates NaN; np.maximum does too
        trellis[t + 1, 1:] = np.maximum(stay, change)
    return trellis


def _tok_emission(frame: np.ndarray, token: int, blank_id: int) -> F32:
    return wildcard_emission(frame, np.array([token]), blank_id)[0]


def backtrack(trellis, emission, tokens, blank_id: int = 0) -> List[Point]:
    """alignment.py:447-481.  Raises AssertionError like the reference when t hits 0 early."""
    t, j = trellis.shape[0] - 1, trellis.shape[1] - 1
    path = [Point(j, t, float(np.exp(emission[t, blank_id])))]
    while j > 0:
        assert t > 0
        p_stay = emission[t - 1, blank_id]
        p_change = _tok_emission(emission[t - 1], tokens[j], blank_id)
        stayed = F32(trellis[t - 1, j] + p_stay)
        changed = F32(trellis[t - 1, j - 1] + p_change)
        t -= 1
        took_change = bool(changed > stayed)
        if took_change:
            j -= 1
        prob = float(np.exp(p_change if took_change else p_stay))
        path.append(Point(j, t, prob))
    while t > 0:
        path.append(Point(j, t - 1, float(np.exp(emission[t - 1, blank_id]))))
        t -= 1
    return path[::-1]


def backtrack_beam(trellis, emission, tokens, blank_id: int = 0, beam_width: int = 5) -> Optional[List[Point]]:
    """alignment.py:500-579.  Candidates are ranked by the predecessor trellis cell only; ties keep
    generation order (Python's stable sort); duplicates are not merged."""
    T, J = trellis.shape[0] - 1, trellis.shape[1] - 1
    beams = [(J, T, trellis[T, J], [Point(J, T, float(np.exp(emission[T, blank_id])))])]
    while beams and beams[0][0] > 0:
        nxt = []
        for (j, t, _score, path) in beams:
            if t <= 0:
                continue
            p_stay = emission[t - 1, blank_id]
            p_change = _tok_emission(emission[t - 1], tokens[j], blank_id)
            stay_score = trellis[t - 1, j]
            change_score = trellis[t - 1, j - 1] if j > 0 else NEG_INF
            if not np.isinf(stay_score):
                nxt.append((j, t - 1, stay_score, path + [Point(j, t - 1, float(np.exp(p_stay)))]))
            if j > 0 and not np.isinf(change_score):
                nxt.append((j - 1, t - 1, change_score, path + [Point(j - 1, t - 1, float(np.exp(p_change)))]))
        beams = sorted(nxt, key=lambda b: b[2], reverse=True)[:beam_width]
        if not beams:
            break
    if not beams:
        return None
    j, t, _s, path = beams[0]
    while t > 0:
        path.append(Point(j, t - 1, float(np.exp(emission[t - 1, blank_id]))))
        t -= 1
    return path[::-1]


def merge_repeats(path: List[Point], transcript: str) -> List[Segment]:
    """alignment.py:597-613."""
    i1 = i2 = 0
    segments = []
    while i1 < len(path):
        while i2 < len(path) and path[i1].token_index == path[i2].token_index:
            i2 += 1
        score = sum(path[k].score for k in range(i1, i2)) / (i2 - i1)
        segments.append(Segment(transcript[path[i1].token_index], path[i1].time_index,
                                path[i2 - 1].time_index + 1, score))
        i1 = i2
    return segments
