"""oracle/vad.py against outputs of the reference's own Binarize / merge_chunks (tests/golden/make_vad_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import vad as ovad

CASES = ["long30", "short_cuts", "silence", "all_speech", "active_edges", "offset_is_none", "shifted_grid"]


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "vad_golden.npz"))


def _unpack(g, name):
    chunk_size, onset, offset, duration, step, start = g[f"{name}_params"]
    offset = None if np.isnan(offset) else float(offset)
    return float(chunk_size), float(onset), offset, float(duration), float(step), float(start)


@pytest.mark.parametrize("name", CASES)
def test_vad_chunks_vs_reference(g, name):
    chunk_size, onset, offset, duration, step, start = _unpack(g, name)
    got = ovad.vad_chunks(g[f"{name}_scores"], duration, step, start, chunk_size, onset, offset)
    want_chunks, want_members = g[f"{name}_chunks"], g[f"{name}_members"]
    assert len(got) == len(want_chunks)
    flat = [(i, s, e) for i, c in enumerate(got) for (s, e) in c["segments"]]
    assert np.array_equal(np.array([[c["start"], c["end"]] for c in got], dtype=np.float64).reshape(-1, 2), want_chunks)
    assert np.array_equal(np.array(flat, dtype=np.float64).reshape(-1, 3), want_members)


def test_silero_merge_vs_reference(g):
    segs = [tuple(r) for r in g["silero_segments"]]
    got = ovad.merge_chunks(segs, 30)
    assert np.array_equal(np.array([[c["start"], c["end"]] for c in got]), g["silero_chunks"])
    flat = [(i, s, e) for i, c in enumerate(got) for (s, e) in c["segments"]]
    assert np.array_equal(np.array(flat), g["silero_members"])


def test_product_merge_chunks_vs_reference(g):
    from whisperx.vads.vad import SegmentX, Vad
    got = Vad.merge_chunks([SegmentX(s, e) for s, e in g["silero_segments"]], 30)
    assert np.array_equal(np.array([[c["start"], c["end"]] for c in got]), g["silero_chunks"])
