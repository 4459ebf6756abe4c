"""
Result writers: txt / vtt / srt / tsv / json / aud files from a transcription result, byte-identical to the
reference's (/root/reference/whisperx/utils.py:170-189 format_timestamp, :192-436 writers and get_writer; options per
/root/reference/whisperx/__main__.py:77-79: max_line_width, max_line_count, highlight_words).  Host-side formatting only; tested
against files produced by the reference's own writers (tests/golden/writers_golden.json).

Subtitles are built in two passes instead of the reference's nested generator: `_layout` assigns every word to a cue and a
line (the line-filling rules of utils.py:236-283), `_cues` turns the groups into (start, end, text) triples (:285-327).
"""
import json
import os
import re
from typing import Callable, Dict, Iterator, List, Optional, TextIO, Tuple

from .utils import LANGUAGES_WITHOUT_SPACES


def format_timestamp(seconds: float, always_include_hours: bool = False, decimal_marker: str = ".") -> str:
    assert seconds >= 0, "non-negative timestamp expected"
    ms = round(seconds * 1000.0)
    h, ms = divmod(ms, 3_600_000)
    m, ms = divmod(ms, 60_000)
    s, ms = divmod(ms, 1_000)
    hours = f"{h:02d}:" if always_include_hours or h > 0 else ""
    return f"{hours}{m:02d}:{s:02d}{decimal_marker}{ms:03d}"


class ResultWriter:
    extension: str

    def __init__(self, output_dir: str):
        self.output_dir = output_dir

    def __call__(self, result: dict, audio_path: str, options: dict):
        stem = os.path.splitext(os.path.basename(audio_path))[0]
        with open(os.path.join(self.output_dir, stem + "." + self.extension), "w", encoding="utf-8") as f:
            self.write_result(result, file=f, options=options)

    def write_result(self, result: dict, file: TextIO, options: dict):
        raise NotImplementedError


def _one_line(text: str) -> str:
    return text.strip().replace("\t", " ")


class WriteTXT(ResultWriter):
    extension = "txt"

    def write_result(self, result: dict, file: TextIO, options: dict):
        for seg in result["segments"]:
            who = seg.get("speaker")
            line = seg["text"].strip()
            print(line if who is None else f"[{who}]: {line}", file=file, flush=True)


class WriteTSV(ResultWriter):
    """start / end in integer milliseconds, tab-separated."""
    extension = "tsv"

    def write_result(self, result: dict, file: TextIO, options: dict):
        print("start\tend\ttext", file=file)
        for seg in result["segments"]:
            print(f"{round(1000 * seg['start'])}\t{round(1000 * seg['end'])}\t{_one_line(seg['text'])}", file=file, flush=True)


class WriteAudacity(ResultWriter):
    """Audacity label track: seconds, tab-separated, no header; the speaker goes in [[ ]] before the text."""
    extension = "aud"

    def write_result(self, result: dict, file: TextIO, options: dict):
        for seg in result["segments"]:
            tag = f"[[{seg['speaker']}]]" if "speaker" in seg else ""
            print(f"{seg['start']}\t{seg['end']}\t{tag}{_one_line(seg['text'])}", file=file, flush=True)


class WriteJSON(ResultWriter):
    extension = "json"

    def write_result(self, result: dict, file: TextIO, options: dict):
        json.dump(result, file, ensure_ascii=False)


Group = Tuple[List[dict], Tuple[float, float, Optional[str]]]


def _layout(segments: List[dict], width_opt: Optional[int], max_lines: Optional[int]) -> Iterator[Group]:
    """Word timings -> cues.  Without both limits every segment is its own cue; with them, words fill lines of at most
    `width` characters, a cue holds at most `max_lines` lines, and a silence of more than 3 s starts a new cue.  A word that
    opens a new line inside a cue carries a leading newline."""
    width = 1000 if width_opt is None else width_opt
    keep_segments = max_lines is None or width_opt is None
    words: List[dict] = []
    owner = None                      # (start, end, speaker) of the segment the cue's first word came from
    used, lines = 0, 1                # characters on the current line, lines in the current cue
    prev_start = segments[0]["start"]
    for seg in segments:
        for k, src in enumerate(seg["words"]):
            w = dict(src)
            timed = "start" in w
            pause = (not keep_segments) and timed and (w["start"] - prev_start > 3.0)
            new_segment = keep_segments and k == 0 and len(words) > 0
            if used > 0 and used + len(w["word"]) <= width and not pause and not new_segment:
                used += len(w["word"])
            else:
                w["word"] = w["word"].strip()
                cue_full = len(words) > 0 and max_lines is not None and (pause or lines >= max_lines)
                if cue_full or new_segment:
                    yield words, owner
                    words, owner, lines = [], None, 1
                elif used > 0:
                    lines += 1
                    w["word"] = "\n" + w["word"]
                used = len(w["word"].strip())
            if owner is None:
                owner = (seg["start"], seg["end"], seg.get("speaker"))
            words.append(w)
            if timed:
                prev_start = w["start"]
    if words:
        yield words, owner


class SubtitlesWriter(ResultWriter):
    always_include_hours: bool
    decimal_marker: str

    def _ts(self, seconds: float) -> str:
        return format_timestamp(seconds, self.always_include_hours, self.decimal_marker)

    def iterate_result(self, result: dict, options: dict) -> Iterator[Tuple[str, str, str]]:
        segments = result["segments"]
        if len(segments) == 0:
            return
        if "words" not in segments[0]:
            for seg in segments:
                text = seg["text"].strip().replace("-->", "->")
                if "speaker" in seg:
                    text = f"[{seg['speaker']}]: {text}"
                yield self._ts(seg["start"]), self._ts(seg["end"]), text
            return
        glue = "" if result["language"] in LANGUAGES_WITHOUT_SPACES else " "
        for words, (seg_start, seg_end, speaker) in _layout(segments, options["max_line_width"], options["max_line_count"]):
            start, end = self._ts(seg_start), self._ts(seg_end)
            text = glue.join(w["word"] for w in words)
            who = "" if speaker is None else f"[{speaker}]: "
            if not (options["highlight_words"] and any("start" in w for w in words)):
                yield start, end, who + text
                continue
            # one cue per timed word with that word underlined, plus the plain text over the gaps between words
            cursor = start
            plain = [w["word"] for w in words]
            for k, w in enumerate(words):
                if "start" not in w:
                    continue
                a, b = self._ts(w["start"]), self._ts(w["end"])
                if cursor != a:
                    yield cursor, a, who + text
                marked = [re.sub(r"^(\s*)(.*)$", r"\1<u>\2</u>", t) if j == k else t for j, t in enumerate(plain)]
                yield a, b, who + " ".join(marked)
                cursor = b


class WriteVTT(SubtitlesWriter):
    extension = "vtt"
    always_include_hours = False
    decimal_marker = "."

    def write_result(self, result: dict, file: TextIO, options: dict):
        print("WEBVTT\n", file=file)
        for start, end, text in self.iterate_result(result, options):
            print(f"{start} --> {end}\n{text}\n", file=file, flush=True)


class WriteSRT(SubtitlesWriter):
    extension = "srt"
    always_include_hours = True
    decimal_marker = ","

    def write_result(self, result: dict, file: TextIO, options: dict):
        for n, (start, end, text) in enumerate(self.iterate_result(result, options), start=1):
            print(f"{n}\n{start} --> {end}\n{text}\n", file=file, flush=True)


_WRITERS = {"txt": WriteTXT, "vtt": WriteVTT, "srt": WriteSRT, "tsv": WriteTSV, "json": WriteJSON}
_OPTIONAL_WRITERS = {"aud": WriteAudacity}


def get_writer(output_format: str, output_dir: str) -> Callable[[dict, str, dict], None]:
    """utils.py:409-436: "all" writes the five standard formats (not the Audacity labels)."""
    if output_format == "all":
        every = [cls(output_dir) for cls in _WRITERS.values()]

        def write_all(result: dict, audio_path: str, options: dict):
            for w in every:
                w(result, audio_path, options)
        return write_all
    if output_format in _OPTIONAL_WRITERS:
        return _OPTIONAL_WRITERS[output_format](output_dir)
    return _WRITERS[output_format](output_dir)
