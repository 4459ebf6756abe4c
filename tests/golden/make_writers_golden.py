"""
Golden files for the result writers, produced by RUNNING THE REFERENCE'S OWN WRITERS
(/root/reference/whisperx/utils.py:192-436, imported in the build container) on seeded results.

Run:  python tests/golden/make_writers_golden.py   ->  tests/golden/writers_golden.json
"""
import io
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import whisperx.utils as ref  # noqa: E402

rnd = random.Random(20261021)
VOCAB = "the quick brown fox jumps over a lazy dog while seventeen ambassadors --> negotiate\ttabs quietly in Zürich 東京 は 晴れ".split(" ")


def make_result(n_seg, with_words, language="en", speakers=False, untimed=0.0, gaps=False):
    t, segs = 0.5, []
    for _ in range(n_seg):
        n = rnd.randint(1, 14)
        words, ws = [], t
        for k in range(n):
            d = rnd.uniform(0.08, 0.6)
            w = {"word": (" " if (k and language == "en" and rnd.random() < 0.5) else "") + rnd.choice(VOCAB)}
            if rnd.random() >= untimed:
                w.update(start=round(ws, 3), end=round(ws + d, 3), score=round(rnd.random(), 3))
            if speakers and rnd.random() < 0.8:
                w["speaker"] = "SPEAKER_%02d" % rnd.randint(0, 2)
            words.append(w)
            ws += d + (rnd.uniform(3.1, 6.0) if (gaps and rnd.random() < 0.15) else rnd.uniform(0.0, 0.2))
        seg = {"start": round(t, 3), "end": round(ws, 3), "text": " " + " ".join(w["word"].strip() for w in words) + " "}
        if with_words:
            seg["words"] = words
        if speakers and rnd.random() < 0.7:
            seg["speaker"] = "SPEAKER_%02d" % rnd.randint(0, 2)
        segs.append(seg)
        t = ws + rnd.uniform(0.1, 2.0)
    res = {"segments": segs, "language": language}
    if with_words:
        res["word_segments"] = [w for s in segs for w in s["words"]]
    return res


cases = []
OPTS = [dict(max_line_width=None, max_line_count=None, highlight_words=False),
        dict(max_line_width=30, max_line_count=2, highlight_words=False),
        dict(max_line_width=18, max_line_count=1, highlight_words=True),
        dict(max_line_width=42, max_line_count=None, highlight_words=True),
        dict(max_line_width=None, max_line_count=3, highlight_words=False)]
RESULTS = [("plain", make_result(6, False)), ("plain_spk", make_result(5, False, speakers=True)), ("empty", {"segments": [], "language": "en"}),
           ("words", make_result(8, True)), ("words_spk", make_result(6, True, speakers=True)), ("words_untimed", make_result(6, True, untimed=0.3)),
           ("words_gaps", make_result(7, True, gaps=True)), ("words_ja", make_result(5, True, language="ja")),
           ("long", {"segments": [{"start": 3599.2, "end": 3725.0049, "text": " over the hour "}], "language": "en"})]
for name, result in RESULTS:
    for oi, options in enumerate(OPTS):
        outs = {}
        for fmt, cls in (("txt", ref.WriteTXT), ("vtt", ref.WriteVTT), ("srt", ref.WriteSRT), ("tsv", ref.WriteTSV), ("json", ref.WriteJSON), ("aud", ref.WriteAudacity)):
            buf = io.StringIO()
            cls("/tmp").write_result(result, file=buf, options=options)
            outs[fmt] = buf.getvalue()
        cases.append({"name": f"{name}/{oi}", "result": name, "options": options, "outputs": outs})
ts = [[s, h, d, ref.format_timestamp(s, h, d)] for s in (0, 0.0004, 0.9996, 59.9995, 61.5, 3599.9999, 3600, 86399.123, 360000.5) for h in (False, True) for d in (".", ",")]
with open(os.path.join(HERE, "writers_golden.json"), "w", encoding="utf-8") as fh:
    json.dump({"results": dict(RESULTS), "cases": cases, "timestamps": ts}, fh, ensure_ascii=False)
print(len(cases), "cases")
