"""
Token-id <-> text for the B200 backend.

No Whisper vocabulary file exists offline, so the default is a deterministic synthetic
detokenizer: every text token id maps to a lowercase pseudo-word (bijective base-26), special tokens
(>= eot) are dropped.  Parity with the reference is defined on token IDs; the text only has to be a
stable function of the ids so that align() has characters to align.

A real vocabulary is loaded with `Tokenizer.from_file(...)`: the rank file Whisper ships (`multilingual.tiktoken` /
`gpt2.tiktoken`: one "base64(token bytes) rank" per line — what mlx_whisper.tokenizer.get_tokenizer reads in the reference,
/root/reference/mlx_whisper_optimized_final.py:278) or a Hugging Face `vocab.json` (GPT-2 byte-to-unicode keys).  Decoding is
the byte-level concatenation + UTF-8 (errors="replace"); encoding is byte-pair merging by rank over the GPT-2 pre-tokeniser
split, needed for `suppress_tokens=[-1]` (the non-speech symbol set) and for SuppressBlank's `encode(" ")`.
"""
import base64
import json
from typing import Callable, Dict, Iterable, List, Optional, Tuple

LANGUAGE_CODES = ("en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt "
                  "la mi ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc "
                  "ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue").split()


_PSEUDO_CACHE: Dict[int, str] = {}


def _pseudo_word(i: int) -> str:
    w = _PSEUDO_CACHE.get(i)
    if w is None:
        w = _PSEUDO_CACHE[i] = _pseudo_word_uncached(i)
    return w


def _pseudo_word_uncached(i: int) -> str:
    s = ""
    i += 26  # at least two letters
    while i > 0:
        i, r = divmod(i, 26)
        s = chr(97 + r) + s
    return s


_GPT2_SPLIT = r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


def _gpt2_byte_decoder() -> Dict[str, int]:
    """GPT-2's printable stand-ins for the 256 byte values (the keys of a Hugging Face vocab.json)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    chars, n = keep[:], 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + n)
            n += 1
    return {chr(c): b for b, c in zip(keep, chars)}


def load_ranks(path: str) -> Dict[bytes, int]:
    """token bytes -> id from a .tiktoken rank file or a Hugging Face vocab.json (entries that are not byte-level text, i.e. the
    special tokens some vocab.json files carry, are skipped)."""
    if path.endswith(".json"):
        dec = _gpt2_byte_decoder()
        with open(path, encoding="utf-8") as fh:
            vocab = json.load(fh)
        out = {}
        for tok, idx in vocab.items():
            if tok.startswith("<|") and tok.endswith("|>"):
                continue  # special tokens are ids >= eot, handled by `specials`
            if all(ch in dec for ch in tok):
                out[bytes(dec[ch] for ch in tok)] = int(idx)
        return out
    out = {}
    with open(path, "rb") as fh:
        for line in fh:
            if line.strip():
                tok, rank = line.split()
                out[base64.b64decode(tok)] = int(rank)
    return out


def _bpe(piece: bytes, ranks: Dict[bytes, int]) -> List[int]:
    """Byte-pair merging by rank (tiktoken's algorithm): repeatedly join the adjacent pair whose concatenation has the lowest rank."""
    parts = [piece[i:i + 1] for i in range(len(piece))]
    while len(parts) > 1:
        best, where = None, -1
        for i in range(len(parts) - 1):
            r = ranks.get(parts[i] + parts[i + 1])
            if r is not None and (best is None or r < best):
                best, where = r, i
        if best is None:
            break
        parts[where:where + 2] = [parts[where] + parts[where + 1]]
    return [ranks[p] for p in parts]


# OpenAI Whisper tokenizer.py `non_speech_tokens`: symbols whose single-token forms (and, for the musical notes, first tokens)
# are suppressed by `suppress_tokens=[-1]` — the reference's default (SURVEY A.3 filter 2)
_NON_SPEECH_SYMBOLS = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』') + \
    "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
_NON_SPEECH_MISC = set("♩♪♫♬♭♮♯")


class Tokenizer:
    @classmethod
    def from_file(cls, path: str, specials: dict, n_vocab: int) -> "Tokenizer":
        """A tokenizer over a real vocabulary file (`*.tiktoken` or vocab.json)."""
        tok = cls(specials, n_vocab)
        tok.ranks = load_ranks(path)
        tok.token_bytes = {i: b for b, i in tok.ranks.items()}
        if max(tok.token_bytes) >= tok.eot:
            raise ValueError(f"{path}: {max(tok.token_bytes) + 1} text tokens do not fit below eot = {tok.eot}")
        return tok

    def encode(self, text: str) -> List[int]:
        """Text -> ids (no special tokens): GPT-2 pre-tokeniser split, then byte-pair merging by rank."""
        if self.ranks is None:
            raise RuntimeError("encode() needs a vocabulary file (Tokenizer.from_file)")
        import regex
        ids: List[int] = []
        for piece in regex.findall(_GPT2_SPLIT, text):
            b = piece.encode("utf-8")
            ids += [self.ranks[b]] if b in self.ranks else _bpe(b, self.ranks)
        return ids

    def non_speech_tokens(self) -> Tuple[int, ...]:
        """What `suppress_tokens=[-1]` stands for (OpenAI Whisper tokenizer.py non_speech_tokens; needs a vocabulary file)."""
        result = {self.encode(" -")[0], self.encode(" '")[0]}
        for symbol in _NON_SPEECH_SYMBOLS + sorted(_NON_SPEECH_MISC):
            for toks in (self.encode(symbol), self.encode(" " + symbol)):
                if len(toks) == 1 or symbol in _NON_SPEECH_MISC:
                    result.add(toks[0])
        return tuple(sorted(result))

    def __init__(self, specials: dict, n_vocab: int, decode_fn: Optional[Callable[[List[int]], str]] = None):
        self.ranks: Optional[Dict[bytes, int]] = None
        self.token_bytes: Optional[Dict[int, bytes]] = None
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
        if self.token_bytes is not None:
            return self.token_bytes.get(int(token), b"").decode("utf-8", errors="replace") if int(token) < self.eot else ""
        return " " + _pseudo_word(int(token)) if 0 <= int(token) < self.eot else ""

    def decode(self, tokens: Iterable[int]) -> str:
        toks = [int(t) for t in tokens if 0 <= int(t) < self.eot]
        if self._decode_fn is not None:
            return self._decode_fn(toks)
        if self.token_bytes is not None:
            return b"".join(self.token_bytes.get(t, b"") for t in toks).decode("utf-8", errors="replace")
        return " ".join(_pseudo_word(t) for t in toks)
