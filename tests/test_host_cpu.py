"""Host-side logic of the product path that needs no GPU: chunk merging, padding, run-length merge,
and the char/word/sentence assembly of align() (fed with oracle paths), against the reference's
golden align() output."""
import json
import os

import numpy as np
import torch

from fake_ctc_model import FakeCTCModel, METADATA, synthetic_speech
from oracle import ctc as octc


def test_merge_chunks_rule():
    from whisperx.vads import SegmentX, Vad
    segs = [SegmentX(0.0, 10.0), SegmentX(11.0, 25.0), SegmentX(26.0, 40.0), SegmentX(41.0, 50.0), SegmentX(90.0, 95.0)]
    out = Vad.merge_chunks(segs, 30, onset=0.5, offset=0.363)
    assert [(c["start"], c["end"]) for c in out] == [(0.0, 25.0), (26.0, 50.0), (90.0, 95.0)]
    assert out[0]["segments"] == [(0.0, 10.0), (11.0, 25.0)]
    one = Vad.merge_chunks([SegmentX(3.0, 50.0)], 30)  # a single over-long region is kept whole
    assert [(c["start"], c["end"]) for c in one] == [(3.0, 50.0)]


def test_pad_or_trim():
    from whisperx.audio import N_SAMPLES, pad_or_trim
    a = np.arange(10, dtype=np.float32)
    assert pad_or_trim(a, 4).tolist() == [0, 1, 2, 3]
    assert pad_or_trim(a, 12).tolist() == list(range(10)) + [0, 0]
    t = torch.arange(6.0).view(2, 3)
    assert pad_or_trim(t, 5, axis=1).shape == (2, 5) and pad_or_trim(t, 2, axis=1).tolist() == [[0, 1], [3, 4]]
    assert pad_or_trim(np.zeros(5, np.float32)).shape == (N_SAMPLES,)


def test_synthetic_cuts():
    from whisperx.vads import synthetic_vad_cuts
    u = synthetic_vad_cuts(1800.0)
    assert len(u) == 60 and u[-1]["end"] == 1800.0
    r = synthetic_vad_cuts(1800.0, mode="ragged")
    assert r[0]["start"] == 0.0 and r[-1]["end"] == 1800.0 and all(c["end"] - c["start"] <= 30.0 + 1e-6 for c in r)


def _oracle_align(transcript, audio, return_chars):
    """align() with the numeric core swapped for the numpy oracle — exercises exactly the host code
    the product runs around kernel K4."""
    import whisperx.alignment as wa

    audio_t = torch.from_numpy(audio)[None]
    model = FakeCTCModel()
    out = []
    for seg in transcript:
        prep = wa._prepare_segment(seg["text"], METADATA["dictionary"], True)
        plain = {"start": seg["start"], "end": seg["end"], "text": seg["text"], "words": [],
                 "chars": [] if return_chars else None}
        if not prep["clean_char"] or seg["start"] >= audio_t.shape[1] / 16000:
            out.append(plain)
            continue
        text_clean = "".join(prep["clean_char"])
        tokens = [METADATA["dictionary"].get(c, -1) for c in text_clean]
        wave = audio_t[:, int(seg["start"] * 16000):int(seg["end"] * 16000)]
        if wave.shape[-1] < 400:
            wave = torch.nn.functional.pad(wave, (0, 400 - wave.shape[-1]))
        em = torch.log_softmax(model(wave)[0], -1)[0].numpy()
        tr = octc.get_trellis(em, tokens, 0)
        path = octc.backtrack_beam(tr, em, tokens, 0, beam_width=2)
        if path is None:
            out.append(plain)
            continue
        pts = [wa.Point(p.token_index, p.time_index, p.score) for p in path]
        cs = wa.merge_repeats(pts, text_clean)
        ratio = (seg["end"] - seg["start"]) * 1 / (tr.shape[0] - 1)
        out += wa._assemble(seg["text"], prep, cs, ratio, seg["start"], True, "nearest", return_chars)
    return out


def _close(a, b, path=""):
    if isinstance(a, dict):
        assert set(a) == set(b), (path, set(a) ^ set(b))
        for k in a:
            _close(a[k], b[k], f"{path}.{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _close(x, y, f"{path}[{i}]")
    elif isinstance(a, (float, np.floating)) or isinstance(b, (float, np.floating)):
        if a is None or b is None:
            assert a is None and b is None, path
        else:
            assert abs(float(a) - float(b)) <= 1.001e-3, (path, a, b)
    else:
        assert a == b, (path, a, b)


def test_align_host_assembly_matches_reference_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "align_golden.json")))
    audio = synthetic_speech(g["audio_seconds"], seed=g["audio_seed"])
    for tag, chars in (("words", False), ("chars", True)):
        got = json.loads(json.dumps(_oracle_align(g["transcript"], audio, chars), default=float))
        _close(got, g["result"][tag]["segments"], tag)


def test_upload_contiguous_run_detection():
    """upload_chunks copies back-to-back views of one float32 array as a single range; anything else (gaps, reordering,
    clipped or strided chunks, other dtypes) must fall back to the per-chunk copy."""
    import numpy as np
    from whisperx.backends.b200 import B200WhisperBackend as B
    a = np.arange(1000, dtype=np.float32)
    run = B._contiguous_run([a[0:300], a[300:600], a[600:1000]], [300, 300, 400])
    assert run is not None and run.shape == (1000,) and np.array_equal(run, a)
    assert np.array_equal(B._contiguous_run([a[100:400]], [300]), a[100:400])
    assert B._contiguous_run([a[0:300], a[301:600]], [300, 299]) is None          # gap
    assert B._contiguous_run([a[300:600], a[0:300]], [300, 300]) is None          # reordered
    assert B._contiguous_run([a[0:300], a[300:600]], [300, 200]) is None          # second chunk clipped
    assert B._contiguous_run([a[::2]], [500]) is None                             # strided
    assert B._contiguous_run([a.astype(np.float64)], [1000]) is None              # other dtype
    assert B._contiguous_run([a.tolist()], [1000]) is None                        # not an array


def test_w2v_weight_layout_round_trip_and_frames():
    """Kernel-layout conversion of the alignment model (weight-norm folded, Q|K|V fused, conv kernels tap-major) inverts
    exactly; the frame-count formula matches torchaudio's own length arithmetic."""
    import torch
    import torchaudio
    from whisperx.align_model import (W2V_BASE_DIMS, frames_for, from_torchaudio_state_dict, kernel_layout_to_torchaudio,
                                      random_init_torchaudio)
    dims = dict(W2V_BASE_DIMS, n_layers=1)
    params = dict(torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params, encoder_num_layers=1)
    m = random_init_torchaudio(params, 0)
    k = from_torchaudio_state_dict(m.state_dict(), dims, "cpu")
    assert k["w2v.pos.w"].shape == (768, 128 * 48) and k["w2v.conv1.w"].shape == (512, 1536) and k["w2v.0.qkv.w"].shape == (2304, 768)
    back = kernel_layout_to_torchaudio(k, dims)
    torch.nn.utils.parametrize.remove_parametrizations(m.encoder.transformer.pos_conv_embed.conv, "weight")
    sd = m.state_dict()
    assert set(back) == set(sd)
    for name, t in sd.items():
        want = t.to(torch.bfloat16).float() if t.dim() >= 2 and "conv_layers.0" not in name else t.float()
        if "pos_conv_embed.conv.weight" in name:  # weight norm folded in fp32 on both sides, then rounded: one bf16 ulp of slack
            assert torch.allclose(back[name], want, rtol=2 ** -7, atol=1e-6), name
        else:
            assert torch.equal(back[name], want), name
    for n in (400, 401, 719, 720, 16000, 123457, 480000):
        _, lens = m.feature_extractor(torch.zeros(1, n), torch.tensor([n]))
        assert frames_for(n) == int(lens[0]) == (n - 400) // 320 + 1
    assert frames_for(399) == 0


def test_align_assembly_numpy_path_equals_pandas_path():
    """align()'s word / sentence assembly: the numpy path taken for single-sentence segments reproduces the reference's
    pandas pipeline value for value (word min / max / rounded nan-mean, dropped NaN rows, char records)."""
    import numpy as np
    from whisperx import alignment as al
    rng = np.random.RandomState(3)
    letters = "abcdefghijklmnopqrstuvwxyz'"
    n_checked = 0
    for case in range(60):
        n_words = int(rng.randint(1, 30))
        words = ["".join(rng.choice(list(letters), size=int(rng.randint(1, 14)))) for _ in range(n_words)]
        if case % 7 == 0:
            words[int(rng.randint(0, n_words))] = "42"          # a word with no character in the dictionary
        text = (" " if case % 3 == 0 else "") + " ".join(words) + ("  " if case % 5 == 0 else "")
        if case == 11:
            text = " 7 8 9 "                                     # nothing alignable at all
        prep = al._prepare_segment(text, {c: i for i, c in enumerate("-|" + letters)}, True)
        prep["clean_cdx"] = [c for c, ch in zip(prep["clean_cdx"], prep["clean_char"]) if ch != "*" or case % 2]
        t = 0
        segs = []
        for _ in prep["clean_cdx"]:
            dur = int(rng.randint(1, 9))
            segs.append(al.Segment("x", t, t + dur, float(rng.rand())))
            t += dur + int(rng.randint(0, 3))
        ratio, t1 = 0.020013342228152101, float(rng.uniform(0, 1800))
        for chars in (False, True):
            runs = ([g.start for g in segs], [g.end for g in segs], [g.score for g in segs])
            a = al._assemble_single_sentence(text, prep, runs, ratio, t1, True, chars)
            b = al._assemble_pandas(text, prep, segs, ratio, t1, True, "nearest", chars)
            assert a == b, (case, a, b)
            n_checked += 1
    assert n_checked == 120
    # long words (8+ characters take ndarray.sum's pairwise order), scores that are multiples of 0.001 (word means then sit ON
    # the rounding boundaries of round(., 3) all the time), double spaces, numpy-array runs as _merge_runs returns them
    for case in range(120):
        n_words = int(rng.randint(1, 40))
        words = ["".join(rng.choice(list(letters + "4"), size=int(rng.randint(1, 20)))) for _ in range(n_words)]
        text = (" " if case % 3 == 0 else "") + (" " if case % 4 else "  ").join(words) + ("  " if case % 5 == 0 else "")
        prep = al._prepare_segment(text, {c: i for i, c in enumerate("-|" + letters)}, True)
        if case % 2:
            prep["clean_cdx"] = [c for c, ch in zip(prep["clean_cdx"], prep["clean_char"]) if ch != "*"]
        t = 0
        segs = []
        for _ in prep["clean_cdx"]:
            dur = int(rng.randint(1, 9))
            segs.append(al.Segment("x", t, t + dur, float(np.round(rng.rand(), 3)) if case % 2 else float(rng.rand())))
            t += dur + int(rng.randint(0, 3))
        runs = (np.array([g.start for g in segs]), np.array([g.end for g in segs]), np.array([g.score for g in segs]))
        t1 = float(rng.uniform(0, 1800))
        for chars in (False, True):
            a = al._assemble(text, prep, runs, 0.020013342228152101, t1, True, "nearest", chars)
            b = al._assemble_pandas(text, prep, segs, 0.020013342228152101, t1, True, "nearest", chars)
            assert a == b, (case, a, b)
    # run-length merge straight from the per-frame arrays == the reference's merge over Point objects
    for T in (1, 7, 300):
        tok = np.sort(rng.randint(0, 40, size=T)).astype(np.int32)
        prob = rng.rand(T).astype(np.float32)
        transcript = "".join(rng.choice(list(letters), size=40))
        a = al._merge_repeats_arrays(tok, prob, transcript)
        b = al.merge_repeats([al.Point(int(tok[t]), t, float(prob[t])) for t in range(T)], transcript)
        assert a == b
