"""The oracle (oracle/*.py) against the golden vectors produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc as octc
from oracle import mel as omel


def _mel_ok(got, ref):
    # acceptance from SURVEY A.1: |d| <= 1e-4 * max(1, |ref|)
    return np.abs(got - ref) <= 1e-4 * np.maximum(1.0, np.abs(ref))


def test_mel_oracle_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_golden.npz"))
    for n_mels, key in ((80, "real_mel80"), (128, "real_mel128")):
        got = omel.log_mel_spectrogram(g["real_audio"], n_mels)
        assert got.shape == g[key].shape
        assert _mel_ok(got, g[key]).all(), np.abs(got - g[key]).max()
    got = omel.log_mel_spectrogram(g["pad_audio"], 128, padding=omel.N_SAMPLES - len(g["pad_audio"]))
    assert got.shape == (128, 3000)
    assert _mel_ok(got[:, g["pad_mel128_frames"]], g["pad_mel128"]).all()
    got = omel.log_mel_spectrogram(g["odd_audio"], 80)
    assert got.shape == g["odd_mel80"].shape and _mel_ok(got, g["odd_mel80"]).all()


def test_mel_per_chunk_max_is_not_batch_global(golden_dir):
    g = np.load(os.path.join(golden_dir, "mel_golden.npz"))
    loud, quiet = g["real_audio"], 1e-3 * g["real_audio"]
    both = omel.log_mel_chunks([loud, quiet], 80, n_samples=48000)
    assert np.array_equal(both[1], omel.log_mel_spectrogram(quiet, 80))


CTC_CASES = ["small", "medium_wild", "peaky", "blank_last", "one_token", "n_eq_t", "n_gt_t", "full"]


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_oracle_bit_exact_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "ctc_golden.npz"))
    em, tokens, blank = g[f"{name}_emission"], g[f"{name}_tokens"].tolist(), int(g[f"{name}_blank"])
    tr = octc.get_trellis(em, tokens, blank)
    if name == "full":
        rows = g["full_trellis_rows"]
        assert np.array_equal(tr[rows], g["full_trellis_sample"])
        fin = np.isfinite(tr)
        assert fin.sum() == g["full_trellis_checksum"][1]
        assert tr[fin].astype(np.float64).sum() == g["full_trellis_checksum"][0]
    else:
        assert np.array_equal(tr, g[f"{name}_trellis"])  # bit-exact incl. +-inf
    if int(g[f"{name}_bt_ok"]):
        p = octc.backtrack(tr, em, tokens, blank)
        assert [q.token_index for q in p] == g[f"{name}_bt_tok"].tolist()
        assert [q.time_index for q in p] == g[f"{name}_bt_time"].tolist()
        np.testing.assert_allclose([q.score for q in p], g[f"{name}_bt_score"], rtol=1e-6)
    else:
        with pytest.raises(AssertionError):
            octc.backtrack(tr, em, tokens, blank)
    p = octc.backtrack_beam(tr, em, tokens, blank, beam_width=2)
    if int(g[f"{name}_beam_ok"]):
        assert [q.token_index for q in p] == g[f"{name}_beam_tok"].tolist()
        assert [q.time_index for q in p] == g[f"{name}_beam_time"].tolist()
        np.testing.assert_allclose([q.score for q in p], g[f"{name}_beam_score"], rtol=1e-6)
    else:
        assert p is None


def test_whisper_oracle_matches_hf_transformers():
    """oracle/whisper.py vs transformers' WhisperModel on a reduced config (same weights)."""
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    from oracle import whisper as ow
    from whisperx.backends.b200_weights import dims_for, from_hf_state_dict

    torch.manual_seed(0)
    cfg = WhisperConfig(vocab_size=300, num_mel_bins=80, d_model=64, encoder_layers=2, decoder_layers=2,
                        encoder_attention_heads=2, decoder_attention_heads=2, encoder_ffn_dim=256,
                        decoder_ffn_dim=256, max_source_positions=1500, max_target_positions=448,
                        activation_function="gelu", scale_embedding=False, pad_token_id=0, bos_token_id=1,
                        eos_token_id=2, decoder_start_token_id=1)
    hf = WhisperForConditionalGeneration(cfg).eval()
    with torch.no_grad():
        for p in hf.parameters():  # HF init is zero-bias: make every parameter matter
            p.add_(0.05 * torch.randn_like(p))
    dims = dims_for("tiny")
    dims.update(n_vocab=300, n_audio_state=64, n_audio_head=2, n_audio_layer=2, n_text_state=64, n_text_head=2,
                n_text_layer=2)
    w = from_hf_state_dict(hf.state_dict())
    mel = torch.randn(2, 80, 3000)
    toks = torch.randint(0, 300, (2, 7))
    with torch.no_grad():
        ref_enc = hf.model.encoder(mel).last_hidden_state
        ref_logits = hf(input_features=mel, decoder_input_ids=toks).logits
        enc = ow.encoder_forward(w, dims, mel)
        cache = ow.DecoderCache(w, dims, enc)
        logits_a = ow.decoder_forward(w, dims, toks[:, :4], cache)   # prompt, then one at a time
        logits_b = torch.cat([ow.decoder_forward(w, dims, toks[:, i:i + 1], cache) for i in range(4, 7)], 1)
    assert torch.allclose(enc, ref_enc, atol=2e-4, rtol=1e-4), (enc - ref_enc).abs().max()
    got = torch.cat([logits_a, logits_b], 1)
    assert torch.allclose(got, ref_logits, atol=5e-4, rtol=1e-4), (got - ref_logits).abs().max()


# ---- greedy decoding rule: the oracle against outputs of the reference's OWN code (tests/golden/make_greedy_golden.py runs
# mlx_whisper_batch_decoder.py / mlx_ultra_optimized_batch.py under a numpy shim of mlx.core) -------------------------------
def test_greedy_update_oracle_vs_reference_golden(golden_dir):
    from oracle import whisper as ow
    g = np.load(os.path.join(golden_dir, "greedy_golden.npz"))
    eot = int(g["update_eot"])
    last = torch.from_numpy(g["update_first_last"])
    sum_lp = torch.zeros(last.shape[0])
    for step in range(g["update_logits"].shape[0]):
        nxt, done, sum_lp = ow.greedy_update(last, torch.from_numpy(g["update_logits"][step]), sum_lp, eot)
        assert nxt.tolist() == g["update_next"][step].tolist(), step
        assert done.tolist() == g["update_done"][step].tolist(), step
        assert np.allclose(sum_lp.numpy(), g["update_sum_logprob"][step], rtol=1e-5, atol=1e-5), step
        last = nxt
    assert bool(g["update_done"][3].all())  # the fixture contains the all-EOT step


def test_greedy_loop_oracle_vs_reference_golden(golden_dir):
    from oracle import whisper as ow
    g = np.load(os.path.join(golden_dir, "greedy_golden.npz"))
    script = torch.from_numpy(g["loop_script"])

    def logits_fn(step, tokens, active):
        lg = script[step].clone()
        lg[~active] = 0.0
        return lg

    toks, sum_lp, nsp, steps = ow.greedy_loop(logits_fn, torch.from_numpy(g["loop_prompt"]), int(g["loop_eot"]),
                                              int(g["loop_no_speech"]), int(g["loop_sample_len"]))
    assert steps == int(g["loop_steps_run"])
    assert toks.tolist() == g["loop_tokens"].tolist()
    assert np.allclose(sum_lp.numpy(), g["loop_sum_logprob"], rtol=1e-5, atol=1e-5)
    assert np.allclose(nsp.numpy(), g["loop_no_speech_prob"], rtol=1e-5, atol=1e-8)


def test_timestamp_clause_oracle_vs_reference_golden(golden_dir):
    """The clause the reference patches (mlx_ultra_optimized_batch.py:38-71): with a history that triggers no other rule
    (two text tokens sampled), oracle.apply_timestamp_rules must suppress text exactly where the reference does."""
    from oracle import whisper as ow
    g = np.load(os.path.join(golden_dir, "greedy_golden.npz"))
    ts_begin = int(g["ts_begin"])
    logits = torch.from_numpy(g["ts_logits"].copy())
    hist = torch.full((logits.shape[0], 2), 7, dtype=torch.long)
    out = ow.apply_timestamp_rules(logits, hist, eot=100, ts_begin=ts_begin, no_timestamps=-1)
    suppressed = torch.isneginf(out[:, :ts_begin]).all(1)
    assert suppressed.tolist() == g["ts_text_suppressed"].tolist()
    assert out.argmax(-1).tolist() == g["ts_argmax"].tolist()


@pytest.mark.parametrize("name", CTC_CASES)
def test_ctc_oracle_beam_widths_vs_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "ctc_golden.npz"))
    gb = np.load(os.path.join(golden_dir, "beam_width_golden.npz"))
    em, tokens, blank = g[f"{name}_emission"], g[f"{name}_tokens"].tolist(), int(g[f"{name}_blank"])
    tr = octc.get_trellis(em, tokens, blank)
    for w in gb["widths"].tolist():
        p = octc.backtrack_beam(tr, em, tokens, blank, beam_width=w)
        if int(gb[f"{name}_w{w}_ok"]):
            assert [q.token_index for q in p] == gb[f"{name}_w{w}_tok"].tolist(), (name, w)
        else:
            assert p is None
