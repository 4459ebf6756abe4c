"""
whisperx (B200-native hot path).  Same lazily-imported top-level API as the reference's
whisperx/__init__.py:9-41 for the functions on the hot path; `load_model(..., backend="b200")`
selects the CUDA backend (the only backend in this package).
"""
import importlib

__all__ = ["load_model", "load_audio", "load_align_model", "align"]


def _lazy(name):
    return importlib.import_module(f"whisperx.{name}")


def load_align_model(*args, **kwargs):
    return _lazy("alignment").load_align_model(*args, **kwargs)


def align(*args, **kwargs):
    return _lazy("alignment").align(*args, **kwargs)


def load_model(*args, **kwargs):
    return _lazy("asr").load_model(*args, **kwargs)


def load_audio(*args, **kwargs):
    return _lazy("audio").load_audio(*args, **kwargs)
