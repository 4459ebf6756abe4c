"""
Token-id <-> text for the B200 backend.

No Whisper vocabulary file exists offline, so the default is a deterministic synthetic
detokenizer: every text token id maps to a lowercase pseudo-word (bijective base-26), special tokens
(>= eot) are dropped.  Parity with the reference is defined on token IDs; the text only has to be a
stable function of the ids so that align() has characters to align.  A real vocabulary can be plugged
in with `Tokenizer(decode_fn=...)`.
"""
from typing import Callable, Iterable, List, Optional

LANGUAGE_CODES = ("en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt "
                  "la mi ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc "
                  "ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue").split()


def _pseudo_word(i: int) -> str:
    s = ""
    i += 26  # at least two letters
    while i > 0:
        i, r = divmod(i, 26)
        s = chr(97 + r) + s
    return s


class Tokenizer:
    def __init__(self, specials: dict, n_vocab: int, decode_fn: Optional[Callable[[List[int]], str]] = None):
        self.specials = dict(specials)
        self.eot = specials["eot"]
        self.sot = specials["sot"]
        self.no_speech = specials["no_speech"]
        self.no_timestamps = specials["no_timestamps"]
        self.transcribe = specials["transcribe"]
        self.translate = specials["translate"]
        self.timestamp_begin = specials["timestamp_begin"]
        self.n_vocab = n_vocab
        self.num_languages = n_vocab - 51765 - 1  # 99 (v1/v2) or 100 (v3)
        self._decode_fn = decode_fn

    def language_token(self, code: str) -> int:
        codes = LANGUAGE_CODES[: self.num_languages]
        if code not in codes:
            raise ValueError(f"unsupported language '{code}'")
        return self.sot + 1 + codes.index(code)

    def language_of(self, token: int) -> str:
        return LANGUAGE_CODES[token - self.sot - 1]

    def prompt(self, language: str = "en", task: str = "transcribe", without_timestamps: bool = True) -> List[int]:
        p = [self.sot, self.language_token(language), self.translate if task == "translate" else self.transcribe]
        if without_timestamps:
            p.append(self.no_timestamps)
        return p

    def decode_piece(self, token: int) -> str:
        """Text of ONE token as it appears inside running text (a leading space marks a word start): what the DTW word
        grouping looks at (mlx_whisper_optimized_final.py:205,213).  Every synthetic pseudo-word is a word of its own."""
        if self._decode_fn is not None:
            return self._decode_fn([int(token)])
        return " " + _pseudo_word(int(token)) if 0 <= int(token) < self.eot else ""

    def decode(self, tokens: Iterable[int]) -> str:
        toks = [int(t) for t in tokens if 0 <= int(t) < self.eot]
        if self._decode_fn is not None:
            return self._decode_fn(toks)
        return " ".join(_pseudo_word(t) for t in toks)
