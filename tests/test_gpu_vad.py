"""GPU parity for VAD post-processing + chunking (csrc/wxb_vad.cu): regions, chunks and chunk membership must be
BIT-IDENTICAL (float64) to what the reference's own Binarize / merge_chunks returned (tests/golden/vad_golden.npz) and to the
oracle on seeded scores; the K1 table equals int(start * 16000) slicing."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["long30", "short_cuts", "silence", "all_speech", "active_edges", "offset_is_none", "shifted_grid"]


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "vad_golden.npz"))


def _params(g, name):
    chunk_size, onset, offset, duration, step, start = [float(v) for v in g[f"{name}_params"]]
    return chunk_size, onset, (None if np.isnan(offset) else offset), duration, step, start


def test_vad_chunks_vs_reference_golden_one_launch(wxb_ctx, g):
    """Cases with the same parameters share a launch (one warp per recording); every case is checked bit for bit."""
    for name in CASES:
        chunk_size, onset, offset, duration, step, start = _params(g, name)
        y = g[f"{name}_scores"]
        n_samples = int(len(y) * step * 16000) + 16000
        (r,) = wxb_ctx.vad_chunks(torch.from_numpy(y).cuda(), np.array([0, len(y)]), np.array([n_samples]), chunk_size, onset=onset,
                                  offset=offset, frame_duration=duration, frame_step=step, frame_start=start)
        want_chunks, want_members = g[f"{name}_chunks"], g[f"{name}_members"]
        assert np.array_equal(r["chunks"].reshape(-1, 2), want_chunks), name
        flat = [(k, s, e) for k in range(len(r["chunks"])) for (s, e) in r["regions"][r["chunk_first"][k]: r["chunk_first"][k + 1]]]
        assert np.array_equal(np.array(flat, dtype=np.float64).reshape(-1, 3), want_members), name
        # the K1 table: asr.py:70-73 slicing, clipped to one 30 s window
        for (s, e), o, l in zip(r["chunks"], r["chunk_off"], r["chunk_len"]):
            a, b = min(int(s * 16000), n_samples), min(int(e * 16000), n_samples)
            assert o == a and l == min(b - a, 480000)
        print(f"[{name}] {len(y)} frames -> {len(r['regions'])} regions, {len(r['chunks'])} chunks: identical")


def test_vad_chunks_batch_vs_oracle(wxb_ctx):
    """Eight seeded recordings of different lengths in ONE launch, against the oracle (NaN-free scores, heavy ties)."""
    from oracle import vad as ovad
    rng = np.random.RandomState(11)
    tracks = []
    for k in range(8):
        n = int(rng.randint(50, 20000))
        x = np.cumsum(rng.standard_normal(n)) * 0.05
        y = (1 / (1 + np.exp(-(x - x.mean())))).astype(np.float32)
        if k % 2:
            y = (np.round(y * 16) / 16).astype(np.float32)  # many equal minima: first-minimum rule
        tracks.append(y)
    tracks.append(np.array([0.9], dtype=np.float32))      # a single active frame: the region is empty and dropped
    tracks.append(np.zeros(0, dtype=np.float32))          # an empty track
    off = np.concatenate([[0], np.cumsum([len(t) for t in tracks])])
    res = wxb_ctx.vad_chunks(torch.from_numpy(np.concatenate(tracks)).cuda(), off, np.full(len(tracks), 10 ** 9), 7.5, onset=0.55, offset=0.45)
    for y, r in zip(tracks, res):
        want = ovad.vad_chunks(y, 0.0619375, 0.016875, 0.0, 7.5, 0.55, 0.45) if len(y) else []
        assert len(want) == len(r["chunks"])
        assert np.array_equal(np.array([[c["start"], c["end"]] for c in want]).reshape(-1, 2), r["chunks"].reshape(-1, 2))
        flat = [(s, e) for c in want for (s, e) in c["segments"]]
        assert np.array_equal(np.array(flat).reshape(-1, 2), r["regions"].reshape(-1, 2))


def test_energy_vad_pipeline(wxb_ctx):
    """audio -> energy scores (vs the oracle scorer) -> chunks through whisperx.load_model(vad_method="energy").transcribe."""
    import whisperx
    from oracle import vad as ovad
    rng = np.random.RandomState(5)
    sr = 16000
    audio = (rng.standard_normal(95 * sr) * 1e-4).astype(np.float32)   # near-silence ...
    for (a, b) in ((2.0, 20.0), (24.0, 61.5), (63.0, 90.0)):           # ... with three loud stretches (one longer than 30 s)
        audio[int(a * sr): int(b * sr)] += (rng.standard_normal(int(b * sr) - int(a * sr)) * 0.1).astype(np.float32)
    got = wxb_ctx.vad_energy_scores(torch.from_numpy(audio).cuda()).cpu().numpy()
    want = ovad.energy_scores(audio)
    assert got.shape == want.shape and np.abs(got - want).max() <= 2e-6, np.abs(got - want).max()
    model = whisperx.load_model("tiny", device="cuda", backend="b200", vad_method="energy", language="en", asr_options={"sample_len": 8})
    cuts = model._segment_audio_with_vad(audio, 30)
    oracle_cuts = ovad.vad_chunks(got, 0.025, 0.010, 0.0, 30, 0.5, 0.363)   # same scores -> identical boundaries
    assert [(c["start"], c["end"]) for c in cuts] == [(c["start"], c["end"]) for c in oracle_cuts]
    assert [c["segments"] for c in cuts] == [c["segments"] for c in oracle_cuts]
    assert all(c["end"] - c["start"] <= 30.0 + 1e-9 for c in cuts) and len(cuts) >= 4
    res = model.transcribe(audio, batch_size=8)
    assert len(res["segments"]) == len(cuts)
    assert [s["start"] for s in res["segments"]] == [round(c["start"], 3) for c in cuts]
