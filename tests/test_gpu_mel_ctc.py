"""GPU parity tests for K1 (log-mel) and K4 (CTC trellis / backtrack / beam-2), through the C-ABI.
Compared against (a) the golden vectors produced by the reference and (b) the oracle on seeded
inputs, incl. BASELINE full sizes."""
import json
import os

import numpy as np
import pytest
import torch

from fake_ctc_model import FakeCTCModel, METADATA, synthetic_speech
from oracle import ctc as octc
from oracle import mel as omel

pytestmark = pytest.mark.gpu

MEL_TOL = 1e-4  # north_star: mel agrees within 1e-4 relative (|d| <= 1e-4*max(1,|ref|))


def _mel_check(got, ref, what):
    got, ref = np.asarray(got, np.float32), np.asarray(ref, np.float32)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    rel_l2 = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err.max() <= MEL_TOL, (what, float(err.max()), float(rel_l2))
    return float(err.max()), float(rel_l2)


def test_logmel_vs_reference_golden(wxb_ctx, golden_dir):
    import whisperx.audio as wa
    g = np.load(os.path.join(golden_dir, "mel_golden.npz"))
    _mel_check(wa.log_mel_spectrogram(g["real_audio"], 80, device="cuda").cpu(), g["real_mel80"], "real80")
    _mel_check(wa.log_mel_spectrogram(g["real_audio"], 128, device="cuda").cpu(), g["real_mel128"], "real128")
    _mel_check(wa.log_mel_spectrogram(g["odd_audio"], 80, device="cuda").cpu(), g["odd_mel80"], "odd80")
    pad = wa.log_mel_spectrogram(g["pad_audio"], 128, padding=wa.N_SAMPLES - len(g["pad_audio"]), device="cuda").cpu().numpy()
    assert pad.shape == (128, 3000)
    _mel_check(pad[:, g["pad_mel128_frames"]], g["pad_mel128"], "pad128")


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_batched_chunks_vs_oracle(wxb_ctx, n_mels):
    import whisperx.audio as wa
    # ragged chunks: full 30 s, short, very short, quiet (per-chunk max must not leak across chunks)
    full = synthetic_speech(30.0, seed=1)
    chunks = [full, synthetic_speech(11.7, seed=2), synthetic_speech(0.31, seed=3), 1e-3 * full[:200000],
              np.zeros(1000, np.float32)]
    got = wa.log_mel_chunks(chunks, n_mels).cpu().numpy()
    ref = omel.log_mel_chunks(chunks, n_mels)
    assert got.shape == (len(chunks), n_mels, 3000)
    for i in range(len(chunks)):
        _mel_check(got[i], ref[i], f"chunk{i}")


CTC_CASES = ["small", "medium_wild", "peaky", "blank_last", "one_token", "n_eq_t", "n_gt_t", "full"]


def _run(ctx, em, tokens, blank, mode, want_trellis=False):
    e = torch.from_numpy(np.ascontiguousarray(em)).cuda()
    tok = torch.from_numpy(np.asarray(tokens, np.int32)).cuda()
    return ctx.ctc_align(e, np.array([0, e.shape[0]]), tok, np.array([0, len(tokens)]), blank, mode, want_trellis)


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_bit_exact_vs_reference_golden(wxb_ctx, golden_dir, name):
    from whisperx._native import CTC_BACKTRACK, CTC_BEAM2, CTC_TRELLIS_ONLY
    g = np.load(os.path.join(golden_dir, "ctc_golden.npz"))
    em, tokens, blank = g[f"{name}_emission"], g[f"{name}_tokens"].tolist(), int(g[f"{name}_blank"])
    T, N = em.shape[0], len(tokens)
    tr = _run(wxb_ctx, em, tokens, blank, CTC_TRELLIS_ONLY, True)["trellis"][:T * N].view(T, N).cpu().numpy()
    ref_tr = octc.get_trellis(em, tokens, blank)  # oracle is pinned bit-exact to the reference (CPU suite)
    assert np.array_equal(tr, ref_tr), "trellis not bit-exact"
    if name != "full":
        assert np.array_equal(tr, g[f"{name}_trellis"])
    r = _run(wxb_ctx, em, tokens, blank, CTC_BACKTRACK)
    if int(g[f"{name}_bt_ok"]):
        assert int(r["status"][0]) == 0
        assert r["path_tok"].cpu().numpy().tolist() == g[f"{name}_bt_tok"].tolist()
        np.testing.assert_allclose(torch.exp(r["path_lp"].cpu()).numpy(), g[f"{name}_bt_score"], rtol=1e-6)
        np.testing.assert_allclose(r["path_prob"].cpu().numpy(), g[f"{name}_bt_score"], rtol=1e-5)
    else:
        assert int(r["status"][0]) == 1
    r = _run(wxb_ctx, em, tokens, blank, CTC_BEAM2)
    if int(g[f"{name}_beam_ok"]):
        assert int(r["status"][0]) == 0
        assert r["path_tok"].cpu().numpy().tolist() == g[f"{name}_beam_tok"].tolist()
        np.testing.assert_allclose(torch.exp(r["path_lp"].cpu()).numpy(), g[f"{name}_beam_score"], rtol=1e-6)
    else:
        assert int(r["status"][0]) == 1


def test_ctc_many_ragged_segments_one_launch(wxb_ctx):
    """BASELINE config 5 shape: many segments (T up to 1499, N ~ U(50,450), 5% wildcards) in ONE
    launch, each bit-exact against the oracle; plus degenerate segments in the same batch."""
    from whisperx._native import CTC_BACKTRACK, CTC_BEAM2
    rng = np.random.RandomState(7)
    V, blank = 29, 0
    shapes = [(1499, 450), (1499, 50), (749, 300), (249, 77), (1, 1), (5, 1), (12, 40), (1499, 201), (600, 599)]
    ems, toks = [], []
    for i, (T, N) in enumerate(shapes):
        g = torch.Generator().manual_seed(100 + i)
        ems.append(torch.log_softmax(torch.randn(T, V, generator=g) / (0.3 if i % 2 else 1.0), -1).numpy())
        t = rng.randint(1, V, size=N)
        t[rng.rand(N) < 0.05] = -1
        toks.append(t.astype(np.int32))
    e = torch.from_numpy(np.concatenate(ems)).cuda()
    tk = torch.from_numpy(np.concatenate(toks)).cuda()
    t_off = np.concatenate([[0], np.cumsum([s[0] for s in shapes])])
    n_off = np.concatenate([[0], np.cumsum([s[1] for s in shapes])])
    for mode, fn in ((CTC_BACKTRACK, "bt"), (CTC_BEAM2, "beam")):
        r = wxb_ctx.ctc_align(e, t_off, tk, n_off, blank, mode)
        status = r["status"].cpu().numpy()
        ptok = r["path_tok"].cpu().numpy()
        plp = r["path_lp"].cpu().numpy()
        for i, (T, N) in enumerate(shapes):
            tr = octc.get_trellis(ems[i], toks[i].tolist(), blank)
            if fn == "bt":
                try:
                    ref = octc.backtrack(tr, ems[i], toks[i].tolist(), blank)
                except AssertionError:
                    ref = None
            else:
                ref = octc.backtrack_beam(tr, ems[i], toks[i].tolist(), blank, beam_width=2)
            if ref is None:
                assert status[i] == 1, (fn, i)
                continue
            assert status[i] == 0, (fn, i)
            a, b = t_off[i], t_off[i + 1]
            assert ptok[a:b].tolist() == [p.token_index for p in ref], (fn, i)
            assert np.allclose(np.exp(plp[a:b]), [p.score for p in ref], rtol=1e-6), (fn, i)


@pytest.mark.parametrize("nmax", [200, 256, 257, 512, 513, 1040, 1088, 1089, 1300])
def test_ctc_every_recurrence_variant_bit_exact(wxb_ctx, nmax):
    """The launch picks the trellis recurrence by its longest segment: the row in registers with 8 / 16 / 34 columns per lane
    (N <= 256 / 512 / 1088), shared-memory rows beyond.  Each variant, at its boundaries, with shorter segments (lanes without
    columns, N = 1, N = 33) and wildcards in the same launch: trellis bits, path indices and probabilities vs the oracle."""
    from whisperx._native import CTC_BACKTRACK, CTC_BEAM2
    rng = np.random.RandomState(nmax)
    V, blank = 29, 0
    shapes = [(1499, nmax), (1499, max(1, nmax // 3)), (400, 33), (300, 1), (1499, nmax - 1)]
    ems, toks = [], []
    for i, (T, N) in enumerate(shapes):
        g = torch.Generator().manual_seed(nmax + i)
        ems.append(torch.log_softmax(torch.randn(T, V, generator=g) / (0.3 if i % 2 else 1.0), -1).numpy())
        t = rng.randint(1, V, size=N)
        if i != 4:
            t[rng.rand(N) < 0.05] = -1
        toks.append(t.astype(np.int32))
    e = torch.from_numpy(np.concatenate(ems)).cuda()
    tk = torch.from_numpy(np.concatenate(toks)).cuda()
    t_off = np.concatenate([[0], np.cumsum([s[0] for s in shapes])])
    n_off = np.concatenate([[0], np.cumsum([s[1] for s in shapes])])
    trs = [octc.get_trellis(ems[i], toks[i].tolist(), blank) for i in range(len(shapes))]
    for mode in (CTC_BACKTRACK, CTC_BEAM2):
        r = wxb_ctx.ctc_align(e, t_off, tk, n_off, blank, mode, want_trellis=True)
        status = r["status"].cpu().numpy()
        ptok = r["path_tok"].cpu().numpy()
        plp = r["path_lp"].cpu().numpy()
        trg = r["trellis"].cpu().numpy()
        o = 0
        for i, (T, N) in enumerate(shapes):
            got = trg[o:o + T * N].reshape(T, N)
            o += T * N
            assert np.array_equal(got.view(np.uint32), np.asarray(trs[i], dtype=np.float32).view(np.uint32)), (nmax, mode, i)
            if mode == CTC_BACKTRACK:
                try:
                    ref = octc.backtrack(trs[i], ems[i], toks[i].tolist(), blank)
                except AssertionError:
                    ref = None
            else:
                ref = octc.backtrack_beam(trs[i], ems[i], toks[i].tolist(), blank, beam_width=2)
            if ref is None:
                assert status[i] == 1, (nmax, mode, i)
                continue
            assert status[i] == 0, (nmax, mode, i)
            a, b = t_off[i], t_off[i + 1]
            assert ptok[a:b].tolist() == [p.token_index for p in ref], (nmax, mode, i)
            assert np.allclose(np.exp(plp[a:b]), [p.score for p in ref], rtol=1e-6), (nmax, mode, i)


def test_log_softmax_rows(wxb_ctx):
    x = torch.randn(1499, 29, generator=torch.Generator().manual_seed(0)) * 4
    got = wxb_ctx.log_softmax_rows_(x.clone().cuda()).cpu()
    assert torch.allclose(got, torch.log_softmax(x, -1), atol=2e-6, rtol=0)


def test_python_seams_match_reference_signatures(wxb_ctx, golden_dir):
    import whisperx.alignment as wa
    g = np.load(os.path.join(golden_dir, "ctc_golden.npz"))
    em, tokens = torch.from_numpy(g["medium_wild_emission"]), g["medium_wild_tokens"].tolist()
    tr = wa.get_trellis(em, tokens, 0)
    assert np.array_equal(tr.cpu().numpy(), g["medium_wild_trellis"])
    p = wa.backtrack(tr, em, tokens, 0)
    assert [q.token_index for q in p] == g["medium_wild_bt_tok"].tolist()
    assert [q.time_index for q in p] == list(range(em.shape[0]))
    p = wa.backtrack_beam(tr, em, tokens, 0, beam_width=2)
    assert [q.token_index for q in p] == g["medium_wild_beam_tok"].tolist()
    assert wa.backtrack_beam(None, torch.from_numpy(g["n_gt_t_emission"]), g["n_gt_t_tokens"].tolist(), 0, beam_width=2) is None
    with pytest.raises(AssertionError):
        wa.backtrack(None, torch.from_numpy(g["n_gt_t_emission"]), g["n_gt_t_tokens"].tolist(), 0)


def test_align_end_to_end_vs_reference_golden(wxb_ctx, golden_dir):
    """whisperx.align() (ours, GPU) returns the same segment / word dicts as the reference's align()
    (golden), given the same emissions (deterministic fake CTC model)."""
    import whisperx
    from test_host_cpu import _close
    g = json.load(open(os.path.join(golden_dir, "align_golden.json")))
    audio = synthetic_speech(g["audio_seconds"], seed=g["audio_seed"])
    for tag, chars in (("words", False), ("chars", True)):
        got = whisperx.align([dict(s) for s in g["transcript"]], FakeCTCModel(), METADATA, audio, "cuda",
                             return_char_alignments=chars)
        got = json.loads(json.dumps(got, default=float))
        _close(got, g["result"][tag], tag)


def test_logmel_features_handoff_equals_f32_path(wxb_ctx):
    """K1 -> K2 hand-off: the bf16 frame-major encoder input written by wxb_logmel_features is the rounded f32 log-mel, and
    encode(None) on it equals encode(mel) bit for bit; ragged / silent chunks included."""
    import whisperx.audio as wa
    from whisperx.backends import b200_weights as bw
    dims = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=128, n_audio_head=2, n_audio_layer=1,
                n_vocab=1000, n_text_ctx=448, n_text_state=128, n_text_head=2, n_text_layer=1)
    wxb_ctx.set_model(dims, bw.to_kernel_layout(bw.init_random_weights(dims, seed=2, std=0.05), dims, "cuda"))
    chunks = [synthetic_speech(30.0, seed=61), synthetic_speech(4.4, seed=62), np.zeros(160000, np.float32), synthetic_speech(30.0, seed=63)[:479999]]
    lens = np.array([len(c) for c in chunks], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int64)
    audio = torch.from_numpy(np.concatenate(chunks)).cuda()
    filt = wa.mel_filters(wxb_ctx.device, 80)
    mel = wxb_ctx.logmel(audio, offs, lens, 480000, 80, filt)
    mel2 = wxb_ctx.logmel_features(audio, offs, lens, 80, filt, want_f32=True)
    assert torch.equal(mel, mel2)
    enc_handoff = wxb_ctx.encode(None, n_chunks=len(chunks))
    enc = wxb_ctx.encode(mel)
    assert torch.equal(enc, enc_handoff)
    with pytest.raises(Exception):
        wxb_ctx.encode(None, n_chunks=len(chunks))  # the hand-off buffer was consumed


@pytest.mark.parametrize("name", ["small", "medium_wild", "peaky", "blank_last", "one_token", "n_eq_t", "n_gt_t", "full"])
def test_backtrack_beam_other_widths_vs_reference_golden(wxb_ctx, golden_dir, name):
    """backtrack_beam for beam widths 1, 3, 5 (the reference's default) and 8: indices bit-exact against paths produced by
    the reference itself (tests/golden/make_beam_golden.py)."""
    import whisperx.alignment as wa
    g = np.load(os.path.join(golden_dir, "ctc_golden.npz"))
    gb = np.load(os.path.join(golden_dir, "beam_width_golden.npz"))
    em = torch.from_numpy(g[f"{name}_emission"])
    tokens = g[f"{name}_tokens"].tolist()
    blank = int(g[f"{name}_blank"])
    for w in gb["widths"].tolist():
        p = wa.backtrack_beam(None, em, tokens, blank, beam_width=w)
        if int(gb[f"{name}_w{w}_ok"]):
            assert p is not None and [q.token_index for q in p] == gb[f"{name}_w{w}_tok"].tolist(), (name, w)
            np.testing.assert_allclose([q.score for q in p], gb[f"{name}_w{w}_score"], rtol=1e-6)
        else:
            assert p is None, (name, w)
    with pytest.raises(NotImplementedError):
        wa.backtrack_beam(None, em, tokens, blank, beam_width=9)
