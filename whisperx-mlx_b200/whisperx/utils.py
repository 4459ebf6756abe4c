"""Small helpers the hot path pulls in from the reference's whisperx/utils.py."""
import zlib

LANGUAGES_WITHOUT_SPACES = ["ja", "zh"]


def exact_div(x: int, y: int) -> int:
    """whisperx/utils.py:141-143."""
    if x % y != 0:
        raise AssertionError(f"{x} is not divisible by {y}")
    return x // y


def compression_ratio(text: str) -> float:
    """whisperx/utils.py:160-162."""
    raw = text.encode("utf-8")
    return len(raw) / len(zlib.compress(raw))


def interpolate_nans(x, method="nearest"):
    """whisperx/utils.py:438-442 — fill NaNs of a pandas Series (used by align())."""
    if x.notnull().sum() > 1:
        return x.interpolate(method=method).ffill().bfill()
    return x.ffill().bfill()
