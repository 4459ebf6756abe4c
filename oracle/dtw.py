"""
ORACLE — test infrastructure only.  Never imported by the product path (whisperx-mlx_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

CPU restatement (numpy) of the reference's single-stage word timing from cross-attention:

    /root/reference/mlx_whisper_optimized_final.py:37-125    CrossAttentionBatchInference: per decode step, the cross-attention
                                                              QK (pre-softmax, already scaled) of every layer, last query row
    /root/reference/mlx_whisper_optimized_final.py:128-253   extract_words_with_dtw: mean over the alignment heads, softmax(10 x),
                                                              median filter 7, per-token normalisation, DTW, word grouping
    /root/reference/median_filter_fix.py:6-22                median_filter_fixed (2-D branch: reflect pad + scipy medfilt per row)

`dtw` itself lives in the un-vendored dependency `mlx-whisper` (mlx_whisper.timing.dtw, pyproject.toml:13, branch pin
`whisperx-optimizations`, no commit): a port of OpenAI whisper/timing.py dtw_cpu + backtrace, restated here from the
published algorithm and cross-checked in tests/test_oracle_cpu.py against the independent implementation in transformers
5.5.0 (`_dynamic_time_warping`).  PINNING: tests/golden/dtw_golden.npz holds outputs of the reference's own
extract_words_with_dtw and median_filter_fixed run in the build container (tests/golden/make_dtw_golden.py, with this
file's dtw standing in for the un-vendored one); the dtw step alone is therefore pinned to the published algorithm and to
transformers, not to reference-run vectors.
"""
from typing import Callable, Dict, List, Sequence

import numpy as np


def median_filter_rows(x: np.ndarray, filter_width: int) -> np.ndarray:
    """median_filter_fix.py:6-22 for 2-D input: reflect-pad each row by width // 2 and take the running median (the crop
    [pad:-pad] never sees scipy's zero padding).  Rows no longer than the pad come back unchanged."""
    pad = filter_width // 2
    if x.shape[-1] <= pad:
        return x
    assert filter_width > 0 and filter_width % 2 == 1
    xp = np.pad(x.astype(np.float32), ((0, 0), (pad, pad)), mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, filter_width, axis=1)
    return np.sort(win, axis=-1)[..., pad].astype(np.float32)


def dtw(x: np.ndarray) -> np.ndarray:
    """OpenAI whisper/timing.py dtw_cpu + backtrace (what mlx_whisper.timing.dtw ports): x f32 [N, M]; returns int [2, P],
    row 0 = indices along N, row 1 = indices along M, monotone path from (0, 0) to (N-1, M-1)."""
    x = np.asarray(x, dtype=np.float32)
    N, M = x.shape
    cost = np.full((N + 1, M + 1), np.inf, dtype=np.float32)
    trace = -np.ones((N + 1, M + 1), dtype=np.int8)
    cost[0, 0] = 0
    for j in range(1, M + 1):
        col_prev = cost[:, j - 1]
        col = cost[:, j]
        xj = x[:, j - 1]
        for i in range(1, N + 1):
            c0, c1, c2 = col_prev[i - 1], col[i - 1], col_prev[i]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            col[i] = xj[i - 1] + c
            trace[i, j] = t
    # backtrace
    i, j = N, M
    trace[0, :] = 2
    trace[:, 0] = 1
    path = []
    while i > 0 or j > 0:
        path.append((i - 1, j - 1))
        t = trace[i, j]
        if t == 0:
            i -= 1
            j -= 1
        elif t == 1:
            i -= 1
        elif t == 2:
            j -= 1
        else:
            raise ValueError("unexpected trace")
    return np.array(path, dtype=np.int64)[::-1, :].T


def alignment_cost(qk_mean: np.ndarray, temperature: float = 10.0, filter_width: int = 7) -> np.ndarray:
    """mlx_whisper_optimized_final.py:182-196: qk_mean f32 [n_tokens, n_frames] (mean over the alignment heads of the
    pre-softmax cross-attention scores) -> the matrix handed (transposed and negated) to dtw: softmax(10 x) over frames,
    median filter 7, zero mean / unit variance per token.  Returns the NORMALISED WEIGHTS [n_tokens, n_frames]
    (dtw runs on -result.T)."""
    w = np.asarray(qk_mean, dtype=np.float32) * np.float32(temperature)
    w = w - w.max(axis=-1, keepdims=True)
    e = np.exp(w)
    w = (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    w = median_filter_rows(w, filter_width)
    mean = w.mean(axis=1, keepdims=True)
    std = w.std(axis=1, keepdims=True) + 1e-8
    return ((w - mean) / std).astype(np.float32)


def group_words(text_tokens: Sequence[int], frames: np.ndarray, decode_one: Callable[[int], str]) -> List[Dict]:
    """mlx_whisper_optimized_final.py:204-253: tokens whose text starts with a space open a new word; a word's start / end
    are alignment[0, first token index] / alignment[0, last token index] x 0.02 s — the reference indexes the PATH by the
    token number (not the path step that reaches the token), which is restated as is."""
    words = []
    strs = [decode_one(t) for t in text_tokens]
    P = len(frames)
    cur, start_idx = "", 0
    for i, s in enumerate(strs):
        if i > 0 and s.startswith(" "):
            if cur.strip():
                sf = frames[start_idx] if start_idx < P else 0
                ef = frames[i - 1] if i - 1 < P else sf
                ef = max(ef, sf)
                words.append({"word": cur.strip(), "start": float(sf * 0.02), "end": float(ef * 0.02), "probability": 1.0})
            cur, start_idx = s, i
        else:
            cur += s
    if cur.strip() and start_idx < P:
        sf = frames[start_idx]
        ef = frames[-1] if P > 0 else sf
        ef = max(ef, sf)
        words.append({"word": cur.strip(), "start": float(sf * 0.02), "end": float(ef * 0.02), "probability": 1.0})
    return words


def extract_words(tokens: Sequence[int], qk_steps: np.ndarray, eot: int, decode_one: Callable[[int], str]):
    """extract_words_with_dtw (:128-253) on qk_steps f32 [n_steps, n_heads, n_frames] = the alignment heads' pre-softmax
    cross-attention row of the LAST query token of every decode forward (step 0 = the prompt forward).  Returns
    (words, alignment [2, P], normalised weights [n_text, n_frames])."""
    text_tokens = [t for t in tokens if t < eot]
    if not text_tokens or len(qk_steps) == 0:
        return [], np.zeros((2, 0), dtype=np.int64), np.zeros((0, 0), dtype=np.float32)
    n = min(len(text_tokens), len(qk_steps))
    qk_mean = np.asarray(qk_steps[:n], dtype=np.float32).mean(axis=1)
    w = alignment_cost(qk_mean)
    alignment = dtw(-w.T)
    return group_words(text_tokens, alignment[0], decode_one), alignment, w
