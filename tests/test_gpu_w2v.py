"""GPU parity for the alignment model: the wav2vec2-base forward on our own kernels, batched over ragged
segments, against torchaudio's module (fp32, CPU) holding exactly the bf16-rounded weights the kernels read; then K4 paths
from our emissions against K4-oracle paths from the oracle's emissions, and whisperx.align() end to end."""
import numpy as np
import pytest
import torch

from fake_ctc_model import DICTIONARY, synthetic_speech
from oracle import ctc as octc

pytestmark = pytest.mark.gpu


def _models(seed=0, head_scale=1.0):
    import torchaudio
    from whisperx.align_model import Wav2Vec2B200, W2V_BASE_DIMS, kernel_layout_to_torchaudio, random_init_torchaudio
    params = torchaudio.pipelines.WAV2VEC2_ASR_BASE_960H._params
    ref = random_init_torchaudio(params, seed)
    sd = ref.state_dict()
    if head_scale != 1.0:
        sd["aux.weight"] = sd["aux.weight"] * head_scale
    ours = Wav2Vec2B200(sd, "cuda", W2V_BASE_DIMS)
    # the oracle holds the bf16-rounded numbers (weight norm folded into a plain conv weight)
    torch.nn.utils.parametrize.remove_parametrizations(ref.encoder.transformer.pos_conv_embed.conv, "weight")
    ref.load_state_dict(kernel_layout_to_torchaudio(ours.kernel_weights, ours.dims))
    return ours, ref.eval()


def test_w2v_emissions_vs_torchaudio_ragged_batch(wxb_ctx):
    from whisperx.align_model import frames_for
    ours, ref = _models()
    waves = [synthetic_speech(30.0, seed=31), synthetic_speech(7.3, seed=32), synthetic_speech(0.9, seed=33), synthetic_speech(12.345, seed=34)]
    emis, t_off = ours.emissions(waves)
    emis = emis.cpu()
    assert t_off.tolist() == [0] + np.cumsum([frames_for(len(w)) for w in waves]).tolist()
    assert frames_for(480000) == 1499
    torch.set_num_threads(16)
    worst = 0.0
    for k, w in enumerate(waves):
        with torch.inference_mode():
            want, _ = ref(torch.from_numpy(w)[None])
        got = emis[t_off[k]:t_off[k + 1]]
        assert got.shape == want[0].shape
        err = (got - want[0]).abs()
        lerr = (torch.log_softmax(got, -1) - torch.log_softmax(want[0], -1)).abs()
        sigma = float(want.std())
        worst = max(worst, float(err.max()) / max(sigma, 1e-6))
        print(f"[w2v segment {k}: {len(w)} samples, T={got.shape[0]}] logits max-abs err {float(err.max()):.4f} (mean {float(err.mean()):.5f}) "
              f"at logit std {sigma:.3f}; log-softmax max-abs err {float(lerr.max()):.4f}")
        assert float(err.max()) <= 0.08 * max(sigma, 1.0) + 0.02
        assert float(err.mean()) <= 0.01 * max(sigma, 1.0) + 0.002
    # a segment's emissions do not depend on what else is in the batch
    alone, _ = ours.emissions([waves[1]])
    assert torch.equal(alone.cpu(), emis[t_off[1]:t_off[2]])


def test_w2v_ctc_paths_vs_oracle(wxb_ctx):
    """K4 (beam-2) on our emissions vs the CPU oracle's trellis / beam on the torch model's emissions, 8 segments of 5-30 s with
    seeded transcripts (5 % wildcards).  Random-init emissions have small decision margins, so bf16-level emission differences
    move some token boundaries by a frame: every divergent segment is listed, the frame-level agreement is gated."""
    from whisperx._native import CTC_BEAM2
    ours, ref = _models(head_scale=6.0)
    rng = np.random.RandomState(7)
    waves = [synthetic_speech(float(s), seed=40 + i) for i, s in enumerate((30.0, 30.0, 21.7, 14.2, 9.9, 5.0, 30.0, 17.3))]
    emis, t_off = ours.emissions(waves)
    wxb_ctx.log_softmax_rows_(emis)
    toks = []
    for k in range(len(waves)):
        T = int(t_off[k + 1] - t_off[k])
        n = int(rng.randint(20, max(21, min(300, T // 2))))
        t = rng.randint(1, 29, size=n).astype(np.int32)
        t[rng.rand(n) < 0.05] = -1
        toks.append(t)
    n_off = np.concatenate([[0], np.cumsum([len(t) for t in toks])]).astype(np.int32)
    res = wxb_ctx.ctc_align(emis, t_off, torch.from_numpy(np.concatenate(toks)).cuda(), n_off, 0, CTC_BEAM2)
    ptok = res["path_tok"].cpu().numpy()
    assert (res["status"].cpu().numpy() == 0).all()
    torch.set_num_threads(16)
    same_seg, frames_same, frames = 0, 0, 0
    for k, w in enumerate(waves):
        with torch.inference_mode():
            want, _ = ref(torch.from_numpy(w)[None])
        e = torch.log_softmax(want[0], -1).numpy()
        path = octc.backtrack_beam(octc.get_trellis(e, toks[k].tolist(), 0), e, toks[k].tolist(), 0, beam_width=2)
        want_tok = np.array([p.token_index for p in path])
        got_tok = ptok[t_off[k]:t_off[k + 1]]
        eq = int((want_tok == got_tok).sum())
        frames_same += eq; frames += len(want_tok)
        same_seg += int(eq == len(want_tok))
        if eq != len(want_tok):
            print(f"DIVERGENCE segment {k}: {len(want_tok) - eq} of {len(want_tok)} frames differ "
                  f"(max boundary shift {int(np.abs(want_tok - got_tok).max())} tokens)")
    print(f"w2v + K4: {same_seg}/{len(waves)} segments with identical paths, frame agreement {frames_same / frames:.4f}")
    assert frames_same / frames >= 0.97


def test_align_with_native_model_matches_torch_model(wxb_ctx):
    """whisperx.align() with the native model (one batched forward) vs the same align() driven by the torch module on the GPU
    (per-segment forward, the reference's structure): same dict structure, word times agree."""
    import whisperx
    ours, ref = _models(head_scale=6.0)
    audio = np.concatenate([synthetic_speech(20.0, seed=50), synthetic_speech(11.0, seed=51)])
    transcript = [{"start": 0.0, "end": 20.0, "text": " hello world this is a test of the aligner"},
                  {"start": 20.0, "end": 31.0, "text": "second segment 2 go"},
                  {"start": 40.0, "end": 45.0, "text": "past the end"}]
    meta = {"language": "en", "dictionary": DICTIONARY, "type": "torchaudio"}
    a = whisperx.align(transcript, ours, meta, audio, "cuda", return_char_alignments=True)
    b = whisperx.align(transcript, ref.cuda(), meta, audio, "cuda", return_char_alignments=True)
    assert [s["text"] for s in a["segments"]] == [s["text"] for s in b["segments"]]
    assert len(a["word_segments"]) == len(b["word_segments"]) > 0
    close = 0
    for wa, wb in zip(a["word_segments"], b["word_segments"]):
        assert wa["word"] == wb["word"] and set(wa) == set(wb)
        if "start" in wa:
            close += int(abs(wa["start"] - wb["start"]) <= 0.045 and abs(wa["end"] - wb["end"]) <= 0.045)
    timed = sum(1 for w in a["word_segments"] if "start" in w)
    print(f"align(): {close}/{timed} word intervals within 45 ms of the torch-model run")
    assert close >= 0.8 * timed
    assert ours.last_stats["segments"] == 2 and ours.last_stats["h2d_bytes"] > 0
