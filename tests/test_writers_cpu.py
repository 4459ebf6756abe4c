"""whisperx/writers.py against files produced by the reference's own writers (tests/golden/make_writers_golden.py).  CPU only."""
import io
import json
import os

import pytest

from whisperx import writers as W


@pytest.fixture(scope="module")
def g(golden_dir):
    with open(os.path.join(golden_dir, "writers_golden.json"), encoding="utf-8") as fh:
        return json.load(fh)


def test_format_timestamp(g):
    for seconds, hours, marker, want in g["timestamps"]:
        assert W.format_timestamp(seconds, hours, marker) == want


def test_every_writer_byte_identical(g):
    classes = {"txt": W.WriteTXT, "vtt": W.WriteVTT, "srt": W.WriteSRT, "tsv": W.WriteTSV, "json": W.WriteJSON, "aud": W.WriteAudacity}
    n = 0
    for case in g["cases"]:
        for fmt, want in case["outputs"].items():
            buf = io.StringIO()
            classes[fmt]("/tmp").write_result(g["results"][case["result"]], file=buf, options=case["options"])
            assert buf.getvalue() == want, (case["name"], fmt)
            n += 1
    assert n == 6 * len(g["cases"])


def test_get_writer_writes_files(g, tmp_path):
    case = next(c for c in g["cases"] if c["name"] == "words/1")
    result = g["results"][case["result"]]
    W.get_writer("all", str(tmp_path))(result, "/some/dir/audio.file.wav", case["options"])
    assert sorted(os.listdir(tmp_path)) == ["audio.file.json", "audio.file.srt", "audio.file.tsv", "audio.file.txt", "audio.file.vtt"]
    for fmt in ("srt", "vtt", "txt", "tsv", "json"):
        assert open(tmp_path / f"audio.file.{fmt}", encoding="utf-8").read() == case["outputs"][fmt]
    W.get_writer("aud", str(tmp_path))(result, "x.mp3", case["options"])
    assert open(tmp_path / "x.aud", encoding="utf-8").read() == case["outputs"]["aud"]
