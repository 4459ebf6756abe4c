"""
Minimal stand-ins for the un-vendored `pyannote.core` (uv.lock: pyannote-core 5.0.0) and `pyannote.audio` names that
/root/reference/whisperx/vads/pyannote.py imports, just enough to RUN the reference's own `Binarize.__call__`
(pyannote.py:134-216) and `Pyannote.merge_chunks` (:282-301) in the build container.  Restated from the published
pyannote.core semantics that code relies on:

  SlidingWindow[i]        Segment(start + i * step, start + i * step + duration)
  Segment.middle          0.5 * (start + end);   bool(Segment) is False when end - start < 1e-6 (an empty segment)
  Annotation[seg, track]  = label: empty segments are not added
  Annotation.get_timeline Timeline of the unique segments sorted by (start, end)

Used only by tests/golden/make_vad_golden.py (test infrastructure; never imported by the product).
"""
import sys
import types

SEGMENT_PRECISION = 1e-6


class Segment:
    def __init__(self, start=0.0, end=0.0):
        self.start, self.end = start, end

    def __bool__(self):
        return bool((self.end - self.start) > SEGMENT_PRECISION)

    @property
    def duration(self):
        return self.end - self.start if self else 0.0

    @property
    def middle(self):
        return 0.5 * (self.start + self.end)

    def __iter__(self):
        yield self.start
        yield self.end

    def __eq__(self, o):
        return (self.start, self.end) == (o.start, o.end)

    def __hash__(self):
        return hash((self.start, self.end))

    def __lt__(self, o):
        return (self.start, self.end) < (o.start, o.end)


class SlidingWindow:
    def __init__(self, duration=0.030, step=0.010, start=0.0, end=None):
        self.duration, self.step, self.start = duration, step, start
        self.end = float("inf") if end is None else end

    def __getitem__(self, i):
        start = self.start + i * self.step
        if start >= self.end:
            return None
        return Segment(start=start, end=start + self.duration)


class SlidingWindowFeature:
    def __init__(self, data, sliding_window, labels=None):
        self.data, self.sliding_window, self.labels = data, sliding_window, labels


class Annotation:
    def __init__(self):
        self._tracks = {}

    def __setitem__(self, key, label):
        segment, track = key
        if not segment:  # empty segments are not added
            return
        self._tracks.setdefault(segment, {})[track] = label

    def __delitem__(self, key):
        segment, track = key
        del self._tracks[segment][track]
        if not self._tracks[segment]:
            del self._tracks[segment]

    def itertracks(self):
        for seg in sorted(self._tracks):
            for track in self._tracks[seg]:
                yield seg, track

    def get_timeline(self):
        return sorted(self._tracks)

    def support(self, collar=0.0):
        raise NotImplementedError("not reached by Pyannote.merge_chunks (pads and min_duration_off are 0)")


def install():
    core = types.ModuleType("pyannote.core")
    core.Segment, core.SlidingWindow, core.SlidingWindowFeature, core.Annotation = Segment, SlidingWindow, SlidingWindowFeature, Annotation
    audio = types.ModuleType("pyannote.audio")
    audio.Model = type("Model", (), {})
    audio.Pipeline = type("Pipeline", (), {})
    io = types.ModuleType("pyannote.audio.core.io")
    io.AudioFile = object
    acore = types.ModuleType("pyannote.audio.core")
    pipes = types.ModuleType("pyannote.audio.pipelines")
    pipes.VoiceActivityDetection = type("VoiceActivityDetection", (), {"__init__": lambda self, *a, **k: None})
    putils = types.ModuleType("pyannote.audio.pipelines.utils")
    putils.PipelineModel = object
    root = types.ModuleType("pyannote")
    root.core, root.audio = core, audio
    sys.modules.update({"pyannote": root, "pyannote.core": core, "pyannote.audio": audio, "pyannote.audio.core": acore,
                        "pyannote.audio.core.io": io, "pyannote.audio.pipelines": pipes, "pyannote.audio.pipelines.utils": putils})
    return core
