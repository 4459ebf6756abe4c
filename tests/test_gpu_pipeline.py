"""The drop-in surface end to end on the GPU: whisperx.load_model(name, backend="b200") ->
model.transcribe(audio, batch_size) -> segment dicts (reference schema, whisperx/types.py)."""
import numpy as np
import pytest
import torch

from fake_ctc_model import FakeCTCModel, METADATA, synthetic_speech

pytestmark = pytest.mark.gpu


def test_load_model_transcribe_schema_and_batch_invariance(wxb_ctx):
    import whisperx
    with pytest.warns(UserWarning):
        model = whisperx.load_model("tiny", device="cuda", backend="b200", language="en", vad_method="uniform",
                                    asr_options={"sample_len": 12})
    audio = synthetic_speech(75.0, seed=8)
    r1 = model.transcribe(audio, batch_size=1)
    r4 = model.transcribe(audio, batch_size=4)
    assert r1["language"] == "en" and len(r1["segments"]) == 3
    for seg in r1["segments"]:
        assert {"text", "start", "end", "tokens", "avg_logprob", "no_speech_prob"} <= set(seg)
        assert isinstance(seg["text"], str) and seg["text"] and seg["end"] > seg["start"]
    assert [s["start"] for s in r1["segments"]] == [0.0, 30.0, 60.0] and r1["segments"][-1]["end"] == 75.0
    # batching must not change a row's numerics
    assert [s["tokens"] for s in r1["segments"]] == [s["tokens"] for s in r4["segments"]]
    # no VAD: the backend cuts 30 s windows itself
    m2 = whisperx.load_model("tiny", device="cuda", backend="b200", language="en", vad_method=None, asr_options={"sample_len": 12})
    r = m2.transcribe(audio, batch_size=8)
    assert [s["tokens"] for s in r["segments"]] == [s["tokens"] for s in r1["segments"]]
    with pytest.raises(ValueError):
        whisperx.load_model("tiny", backend="mlx_lightning")
    assert isinstance(model.detect_language(audio[:160000]), str)


def test_word_timestamps_through_pipeline(wxb_ctx):
    import whisperx
    model = whisperx.load_model("tiny", device="cuda", backend="b200", language="en", vad_method="uniform",
                                asr_options={"sample_len": 10}, align_model=(FakeCTCModel(), METADATA))
    audio = synthetic_speech(42.0, seed=9)
    r = model.transcribe(audio, batch_size=2, word_timestamps=True)
    assert set(r) >= {"segments", "word_segments", "language"}
    assert len(r["word_segments"]) > 0
    w = r["word_segments"][0]
    assert {"word", "start", "end", "score"} <= set(w) and 0.0 <= w["start"] <= w["end"] <= 42.0
