"""
Generates the committed golden fixtures by RUNNING THE REFERENCE ITSELF (imported from
/root/reference, which only exists in the build container; the fixtures travel, the reference
does not).  Run:  python tests/golden/make_golden.py

  mel_golden.npz    whisperx.audio.log_mel_spectrogram        (audio.py:112-159)
  ctc_golden.npz    whisperx.alignment.get_trellis / backtrack / backtrack_beam(beam_width=2)
  align_golden.json whisperx.alignment.align() end to end on a deterministic fake CTC model

nltk is not installed here; the reference's only use of it (Punkt sentence spans,
alignment.py:25,190-194) is satisfied by a stub that returns one span per text — the product's
align() uses the same one-span rule when nltk is absent.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, "/root/reference")

# ---- nltk stub ------------------------------------------------------------------------------
nltk = types.ModuleType("nltk")
tok = types.ModuleType("nltk.tokenize")
punkt = types.ModuleType("nltk.tokenize.punkt")


class PunktParameters:
    abbrev_types = set()


class PunktSentenceTokenizer:
    def __init__(self, params=None):
        pass

    def span_tokenize(self, text):
        return [(0, len(text))]


punkt.PunktParameters, punkt.PunktSentenceTokenizer = PunktParameters, PunktSentenceTokenizer
nltk.tokenize, tok.punkt = tok, punkt
sys.modules.update({"nltk": nltk, "nltk.tokenize": tok, "nltk.tokenize.punkt": punkt})

import whisperx.audio as ref_audio  # noqa: E402
import whisperx.alignment as ref_align  # noqa: E402
from fake_ctc_model import FakeCTCModel, METADATA, synthetic_speech  # noqa: E402

torch.set_num_threads(1)


def make_mel():
    out = {}
    real = np.load("/root/reference/audio_sample.npy").astype(np.float32)  # 5 s of real speech
    out["real_audio"] = real[:48000]
    out["real_mel80"] = ref_audio.log_mel_spectrogram(real[:48000], 80).numpy()
    out["real_mel128"] = ref_audio.log_mel_spectrogram(real[:48000], 128).numpy()
    # hot-path convention: a short chunk zero-padded to 30 s; keep every 13th frame + the tail
    syn = synthetic_speech(7.3, seed=11)
    full = ref_audio.log_mel_spectrogram(syn, 128, padding=ref_audio.N_SAMPLES - len(syn)).numpy()
    out["pad_audio"] = syn
    out["pad_mel128_frames"] = np.concatenate([np.arange(0, 3000, 13), np.arange(2990, 3000)]).astype(np.int32)
    out["pad_mel128"] = full[:, out["pad_mel128_frames"]]
    # length that is not a multiple of the hop, no padding
    odd = synthetic_speech(1.0, seed=5)[:15987]
    out["odd_audio"] = odd
    out["odd_mel80"] = ref_audio.log_mel_spectrogram(odd, 80).numpy()
    np.savez_compressed(os.path.join(HERE, "mel_golden.npz"), **out)
    print("mel:", {k: v.shape for k, v in out.items()})


def rand_emission(T, V, seed, temp=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.log_softmax(torch.randn(T, V, generator=g) / temp, dim=-1)


def path_arrays(path):
    return (np.array([p.token_index for p in path], np.int32), np.array([p.time_index for p in path], np.int32),
            np.array([p.score for p in path], np.float64))


def make_ctc():
    cases = [  # name, T, N, V, blank, wildcard fraction, temperature
        ("small", 50, 7, 29, 0, 0.0, 1.0),
        ("medium_wild", 249, 60, 29, 0, 0.1, 1.0),
        ("peaky", 180, 40, 29, 0, 0.05, 0.2),
        ("blank_last", 120, 25, 32, 31, 0.05, 1.0),
        ("one_token", 40, 1, 29, 0, 0.0, 1.0),
        ("n_eq_t", 30, 30, 29, 0, 0.0, 1.0),
        ("n_gt_t", 20, 26, 29, 0, 0.0, 1.0),
        ("full", 1499, 450, 29, 0, 0.05, 0.5),
    ]
    out = {"names": np.array([c[0] for c in cases])}
    for name, T, N, V, blank, wild, temp in cases:
        seed = abs(hash(name)) % 100000 if False else sum(map(ord, name))
        em = rand_emission(T, V, seed, temp)
        rng = np.random.RandomState(seed)
        labels = [v for v in range(V) if v != blank]
        tokens = [int(rng.choice(labels)) for _ in range(N)]
        tokens = [-1 if rng.rand() < wild else t for t in tokens]
        trellis = ref_align.get_trellis(em, tokens, blank)
        out[f"{name}_emission"] = em.numpy()
        out[f"{name}_tokens"] = np.array(tokens, np.int32)
        out[f"{name}_blank"] = np.array(blank)
        if name == "full":  # 2.7 MB trellis: keep a strided sample + fp64 checksum
            tr = trellis.numpy()
            out[f"{name}_trellis_rows"] = np.arange(0, T, 37).astype(np.int32)
            out[f"{name}_trellis_sample"] = tr[out[f"{name}_trellis_rows"]]
            finite = np.isfinite(tr)
            out[f"{name}_trellis_checksum"] = np.array([tr[finite].astype(np.float64).sum(), finite.sum()])
        else:
            out[f"{name}_trellis"] = trellis.numpy()
        try:
            p = ref_align.backtrack(trellis, em, tokens, blank)
            out[f"{name}_bt_tok"], out[f"{name}_bt_time"], out[f"{name}_bt_score"] = path_arrays(p)
            out[f"{name}_bt_ok"] = np.array(1)
        except AssertionError:
            out[f"{name}_bt_ok"] = np.array(0)
        p = ref_align.backtrack_beam(trellis, em, tokens, blank, beam_width=2)
        if p is None:
            out[f"{name}_beam_ok"] = np.array(0)
        else:
            out[f"{name}_beam_ok"] = np.array(1)
            out[f"{name}_beam_tok"], out[f"{name}_beam_time"], out[f"{name}_beam_score"] = path_arrays(p)
        print(name, "bt_ok", int(out[f"{name}_bt_ok"]), "beam_ok", int(out[f"{name}_beam_ok"]))
    np.savez_compressed(os.path.join(HERE, "ctc_golden.npz"), **out)


ALIGN_TRANSCRIPT = [
    {"start": 0.5, "end": 4.25, "text": " Hello there, this is a test of the aligner."},
    {"start": 4.25, "end": 9.0, "text": "It costs 25 dollars in 2024, doesn't it?"},
    {"start": 9.0, "end": 9.02, "text": "ok"},
    {"start": 9.5, "end": 12.0, "text": "   "},
    {"start": 10.0, "end": 11.98, "text": "Mr. Smith went to Washington. He stayed there!  "},
    {"start": 13.0, "end": 14.0, "text": "beyond the end"},
    {"start": 11.2, "end": 11.9, "text": "supercalifragilisticexpialidocious antidisestablishmentarianism pneumonoultramicroscopic"},
]


def make_align():
    audio = synthetic_speech(12.0, seed=3)
    res = {}
    for tag, chars in (("words", False), ("chars", True)):
        r = ref_align.align([dict(s) for s in ALIGN_TRANSCRIPT], FakeCTCModel(), METADATA, audio, "cpu",
                            return_char_alignments=chars)
        res[tag] = json.loads(json.dumps(r, default=float))
    with open(os.path.join(HERE, "align_golden.json"), "w") as f:
        json.dump({"transcript": ALIGN_TRANSCRIPT, "audio_seconds": 12.0, "audio_seed": 3, "result": res}, f, indent=1)
    print("align: segments", len(res["words"]["segments"]), "words", len(res["words"]["word_segments"]))


if __name__ == "__main__":
    make_mel()
    make_ctc()
    make_align()
