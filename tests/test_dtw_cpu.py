"""oracle/dtw.py against vectors produced by the reference's own code (tests/golden/make_dtw_golden.py) and, for the
un-vendored dtw step, against the independent implementation shipped in transformers.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import dtw as odtw


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "dtw_golden.npz"))


def test_median_filter_vs_reference(g):
    assert np.array_equal(odtw.median_filter_rows(g["medfilt_in"], 7), g["medfilt_w7"])
    assert np.array_equal(odtw.median_filter_rows(g["medfilt_in"], 3), g["medfilt_w3"])
    assert np.array_equal(odtw.median_filter_rows(g["medfilt_short_in"], 7), g["medfilt_short_w7"])


def _decode_one(t):  # the golden script's tokenizer
    return (" " if (t % 3 == 0) else "") + "w%d" % t


@pytest.mark.parametrize("name", ["a", "long", "flat", "one", "none"])
def test_extract_words_vs_reference(g, name):
    tokens = g[f"words_{name}_tokens"].tolist()
    qk = g[f"words_{name}_qk"].astype(np.float32)
    want = json.loads(bytes(g[f"words_{name}_json"]).decode())
    words, alignment, w = odtw.extract_words(tokens, qk, 1000, _decode_one)
    assert words == want
    if f"words_{name}_cost" in g.files:
        # the matrix the reference handed to dtw (its softmax / median filter / normalisation, run under the numpy shim)
        np.testing.assert_allclose(-w.T, g[f"words_{name}_cost"], rtol=0, atol=2e-5)
        assert np.array_equal(alignment, g[f"words_{name}_path"])


def test_dtw_vs_transformers():
    """mlx_whisper.timing.dtw is un-vendored: the restatement is cross-checked against transformers' numpy DTW."""
    from transformers.models.whisper.generation_whisper import _dynamic_time_warping
    rng = np.random.RandomState(7)
    for (N, M) in ((40, 9), (9, 40), (1, 5), (5, 1), (120, 31)):
        x = rng.standard_normal((N, M)).astype(np.float32)
        if N == 40:
            x[10:20, 3] = x[10, 3]  # equal costs: the tie rules decide
        p = odtw.dtw(x)
        ti, tj = _dynamic_time_warping(x.astype(np.float64))
        assert np.array_equal(p[0], ti) and np.array_equal(p[1], tj)
        assert p[0, 0] == 0 and p[1, 0] == 0 and p[0, -1] == N - 1 and p[1, -1] == M - 1
        assert (np.diff(p[0]) >= 0).all() and (np.diff(p[1]) >= 0).all()


@pytest.mark.parametrize("name", ["a", "long", "flat", "one"])
def test_product_word_grouping_vs_reference(g, name):
    """The host half of the product path (whisperx/word_timing.py) on the reference's own path -> the reference's words."""
    from whisperx.word_timing import words_from_path
    tokens = [int(t) for t in g[f"words_{name}_tokens"] if t < 1000]
    want = json.loads(bytes(g[f"words_{name}_json"]).decode())
    assert words_from_path(tokens, g[f"words_{name}_path"][0], _decode_one) == want
    shifted = words_from_path(tokens, g[f"words_{name}_path"][0], _decode_one, offset=60.0)
    assert [w["word"] for w in shifted] == [w["word"] for w in want]
    assert all(abs(a["start"] - b["start"] - 60.0) < 1e-9 for a, b in zip(shifted, want))


def test_alignment_head_tables():
    from whisperx.word_timing import alignment_heads
    assert alignment_heads("large-v3", 32, 20) == [[7, 0], [10, 17], [12, 18], [13, 12], [16, 1], [17, 14], [19, 11], [21, 4], [24, 1], [25, 6]]
    assert len(alignment_heads("tiny", 4, 6)) == 6 and len(alignment_heads("large-v3-turbo", 4, 20)) == 6
    assert alignment_heads("unknown", 4, 2) == [[2, 0], [2, 1], [3, 0], [3, 1]]  # upstream default: upper half of the decoder
