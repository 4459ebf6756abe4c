"""Backends of this package: only the B200 CUDA backend (imported lazily — no device import at
package import time, unlike the reference's eager `from .mlx_whisper import ...`)."""

__all__ = ["WhisperBackend", "B200WhisperBackend"]


def __getattr__(name):
    if name == "WhisperBackend":
        from .base import WhisperBackend
        return WhisperBackend
    if name == "B200WhisperBackend":
        from .b200 import B200WhisperBackend
        return B200WhisperBackend
    raise AttributeError(name)
