"""
Single-stage word timestamps from the decoder's cross-attention: the host half of
/root/reference/mlx_whisper_optimized_final.py:128-253 (extract_words_with_dtw).  The numeric half — alignment-head
scores, softmax(10 x), median filter 7, normalisation and the DTW itself — runs on the GPU (csrc/wxb_dtw.cu) from the
queries the decode kernel logged; what is left here is the token -> word grouping and the seconds conversion, kept
operation for operation so the returned dicts equal the reference's on the same path.
"""
import base64
import gzip
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

# Published OpenAI-Whisper alignment-head masks (whisper/__init__.py `_ALIGNMENT_HEADS`: base85(gzip(bool[n_layers, n_heads]))),
# what `model.alignment_heads` holds in the reference (mlx_whisper_optimized_final.py:147).
_ALIGNMENT_HEADS = {
    "tiny.en": (b"ABzY8J1N>@0{>%R00Bk>$p{7v037`oCl~+#00", 4, 6),
    "tiny": (b"ABzY8bu8Lr0{>%RKn9Fp%m@SkK7Kt=7ytkO", 4, 6),
    "base": (b"ABzY8KQ!870{>%RzyTQH3`Q^yNP!>##QT-<FaQ7m", 6, 8),
    "small": (b"ABzY8DmU8=0{>%Rpa?J`kvJ6qF(V^F86#Xh7JUGMK}P<N0000", 12, 12),
    "medium": (b"ABzY8B0Jh+0{>%R7}kK1fFL7w6%<-Pf*t^=N)Qr&0RR9", 24, 16),
    "large-v2": (b"ABzY8zd+h!0{>%R7=D0pU<_bnWW*tkYAhobTNnu$jnkEkXqp)j;w1Tzk)UH3X%SZd&fFZ2fC2yj", 32, 20),
    "large-v3": (b"ABzY8gWO1E0{>%R7(9S+Kn!D~%ngiGaR?*L!iJG9p-nab0JQ=-{D1-g00", 32, 20),
    "large-v3-turbo": (b"ABzY8j^C+e0{>%RARaKHP%t(lGR*)0g!tONPyhe`", 4, 20),
}
MAX_HEADS = 127  # wxb_decode_collect_heads


def alignment_heads(model_name: str, n_text_layer: int, n_text_head: int) -> List[List[int]]:
    """[(layer, head), ...] of the named model; unknown names fall back to upstream's default (every head of the upper
    half of the decoder), capped at the kernel's table size."""
    entry = _ALIGNMENT_HEADS.get(model_name)
    if entry is not None and entry[1] == n_text_layer and entry[2] == n_text_head:
        mask = np.frombuffer(gzip.decompress(base64.b85decode(entry[0])), dtype=bool).reshape(entry[1], entry[2])
        return [[int(l), int(h)] for l, h in np.argwhere(mask)]
    heads = [[l, h] for l in range(n_text_layer // 2, n_text_layer) for h in range(n_text_head)]
    return heads[-MAX_HEADS:]


def words_from_path(text_tokens: Sequence[int], path_frames: np.ndarray, decode_one: Callable[[int], str],
                    offset: float = 0.0) -> List[Dict]:
    """mlx_whisper_optimized_final.py:204-253.  A token whose text starts with a space opens a new word; the word's start /
    end are `alignment[0, k]` at k = index of its first / last TOKEN (the reference indexes the path by token number), in
    0.02 s frames; `offset` is the chunk's position in the recording (:444-449)."""
    frames = np.asarray(path_frames)
    P = len(frames)
    pieces = [decode_one(int(t)) for t in text_tokens]
    words: List[Dict] = []

    def emit(word: str, first: int, last_frame):
        sf = frames[first] if first < P else 0
        ef = max(last_frame(sf), sf)
        words.append({"word": word.strip(), "start": float(sf * 0.02) + offset, "end": float(ef * 0.02) + offset, "probability": 1.0})

    cur, first = "", 0
    for i, piece in enumerate(pieces):
        if i > 0 and piece.startswith(" "):
            if cur.strip():
                emit(cur, first, lambda sf, i=i: frames[i - 1] if i - 1 < P else sf)
            cur, first = piece, i
        else:
            cur += piece
    if cur.strip() and first < P:
        emit(cur, first, lambda sf: frames[-1] if P > 0 else sf)
    return words


def dtw_word_timestamps(ctx, tokens_per_seq: Sequence[Sequence[int]], eot: int, prompt_len: int, decode_one: Callable[[int], str],
                        offsets: Optional[Sequence[float]] = None, n_frames: int = 1500, return_paths: bool = False):
    """Word lists of every sequence of the LAST decode on `ctx` (which must have run with alignment heads selected,
    Context.collect_alignment_heads): three launches for the whole batch — scores from the logged queries and the resident
    cross-K cache, cost rows, one DTW CTA per sequence — then the host grouping."""
    text = [[int(t) for t in toks if int(t) < eot] for toks in tokens_per_seq]
    # rows = decode steps 0 .. n_text-1 (mlx_whisper_optimized_final.py:157,176-179): the forward at position prompt_len-1+s
    n_rows = np.array([len(t) for t in text], dtype=np.int32)
    if int(n_rows.sum()) == 0:
        empty = [[] for _ in text]
        return (empty, [np.zeros((2, 0), dtype=np.int32) for _ in text]) if return_paths else empty
    qk = ctx.dtw_scores(n_rows, prompt_len - 1, n_frames)
    cost = ctx.dtw_cost(qk, temperature=10.0, medfilt_width=7)
    paths = ctx.dtw_path(cost, n_rows)
    out = []
    for b, toks in enumerate(text):
        off = float(offsets[b]) if offsets is not None else 0.0
        out.append(words_from_path(toks, paths[b][0], decode_one, off) if toks else [])
    return (out, paths) if return_paths else out
