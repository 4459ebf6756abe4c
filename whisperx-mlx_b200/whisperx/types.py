"""Result schemas of the reference API (whisperx/types.py:4-69), restated as TypedDicts."""
from typing import List, Optional, Tuple, TypedDict


class SingleWordSegment(TypedDict, total=False):
    word: str
    start: float
    end: float
    score: float


class SingleCharSegment(TypedDict, total=False):
    char: str
    start: float
    end: float
    score: float


class SingleSegment(TypedDict):
    start: float
    end: float
    text: str


class SegmentData(TypedDict):
    clean_char: List[str]
    clean_cdx: List[int]
    clean_wdx: List[int]
    sentence_spans: List[Tuple[int, int]]


class SingleAlignedSegment(TypedDict):
    start: float
    end: float
    text: str
    words: List[SingleWordSegment]
    chars: Optional[List[SingleCharSegment]]


class TranscriptionResult(TypedDict):
    segments: List[SingleSegment]
    language: str


class AlignedTranscriptionResult(TypedDict):
    segments: List[SingleAlignedSegment]
    word_segments: List[SingleWordSegment]
