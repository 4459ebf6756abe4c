"""Real-vocabulary tokenizer: rank-file / vocab.json loading, byte-level decode, BPE encode — cross-checked
against the tiktoken package on a synthetic vocabulary (no Whisper vocabulary file exists offline).  CPU only."""
import base64
import json

import numpy as np
import pytest

from whisperx.backends import b200_weights as bw
from whisperx.tokenizer import _GPT2_SPLIT, Tokenizer, _gpt2_byte_decoder


def _synthetic_ranks():
    """256 byte tokens + merges learnt greedily from a small corpus (a valid BPE vocabulary: every merge joins two older tokens)."""
    corpus = ("the quick brown fox jumps over the lazy dog; then the other thing happened -- (really) \"quoted\" [x] {y} "
              "♪♪ music ♪ héllo wörld 你好 世界 12345 'tis the season's they're we've I'm he'll she'd ").encode("utf-8") * 3
    ranks = {bytes([i]): i for i in range(256)}
    words = [list(bytes([b]) for b in w) for w in corpus.split(b" ")]
    words = [[b" "] + w if i else w for i, w in enumerate(words)]
    for _ in range(150):
        pairs = {}
        for w in words:
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] = pairs.get((a, b), 0) + 1
        if not pairs:
            break
        (a, b), n = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        if n < 2:
            break
        ranks[a + b] = len(ranks)
        for w in words:
            i = 0
            while i < len(w) - 1:
                if w[i] == a and w[i + 1] == b:
                    w[i:i + 2] = [a + b]
                else:
                    i += 1
    return ranks


@pytest.fixture(scope="module")
def vocab_files(tmp_path_factory):
    d = tmp_path_factory.mktemp("vocab")
    ranks = _synthetic_ranks()
    tk = d / "multilingual.tiktoken"
    with open(tk, "wb") as fh:
        for tok, r in sorted(ranks.items(), key=lambda kv: kv[1]):
            fh.write(base64.b64encode(tok) + b" " + str(r).encode() + b"\n")
    enc = {v: k for k, v in _gpt2_byte_decoder().items()}
    hf = d / "vocab.json"
    with open(hf, "w", encoding="utf-8") as fh:
        vocab = {"".join(enc[b] for b in tok): r for tok, r in ranks.items()}
        vocab["<|endoftext|>"] = 50257
        json.dump(vocab, fh, ensure_ascii=False)
    return ranks, str(tk), str(hf)


def test_vocab_files_and_tiktoken_agreement(vocab_files):
    import tiktoken
    ranks, tk, hf = vocab_files
    sp = bw.special_tokens(bw.dims_for("large-v3"))
    a = Tokenizer.from_file(tk, sp, 51866)
    b = Tokenizer.from_file(hf, sp, 51866)
    assert a.ranks == ranks and b.ranks == ranks
    ref = tiktoken.Encoding("synthetic", pat_str=_GPT2_SPLIT, mergeable_ranks=ranks, special_tokens={"<|endoftext|>": 50257})
    for text in ("the quick brown fox", " then they're here -- (ok)", "héllo 你好 ♪♪ 12345", "", " ", "unseen zzzqqq \t tabs\n\nnew"):
        ids = a.encode(text)
        assert ids == ref.encode(text), text
        assert a.decode(ids) == text and b.decode(ids) == text
        assert "".join(a.decode_piece(t) for t in ids) == text or "�" in "".join(a.decode_piece(t) for t in ids)
    # special tokens and out-of-range ids are dropped by decode, like the synthetic tokenizer does
    assert a.decode(a.encode("the fox") + [sp["eot"], sp["timestamp_begin"] + 3]) == "the fox"
    # a multi-byte character split over tokens decodes with the replacement character per piece but correctly as a whole
    ids = a.encode("你")
    assert a.decode(ids) == "你"


def test_non_speech_tokens_and_backend_options(vocab_files):
    import tiktoken
    ranks, tk, _ = vocab_files
    sp = bw.special_tokens(bw.dims_for("tiny"))
    tok = Tokenizer.from_file(tk, sp, 51865)
    ref = tiktoken.Encoding("synthetic", pat_str=_GPT2_SPLIT, mergeable_ranks=ranks, special_tokens={})
    # OpenAI tokenizer.py non_speech_tokens, evaluated with tiktoken as the encoder
    symbols = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』') + "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
    misc = set("♩♪♫♬♭♮♯")
    want = {ref.encode(" -")[0], ref.encode(" '")[0]}
    for s_ in symbols + list(misc):
        for t in (ref.encode(s_), ref.encode(" " + s_)):
            if len(t) == 1 or s_ in misc:
                want.add(t[0])
    assert tok.non_speech_tokens() == tuple(sorted(want))
    assert tok.encode(" ") == ref.encode(" ")


def test_language_tokens_and_prompt():
    sp = bw.special_tokens(bw.dims_for("large-v3"))
    tok = Tokenizer(sp, 51866)
    assert tok.num_languages == 100 and tok.language_token("en") == 50259 and tok.language_token("yue") == 50358
    assert tok.prompt("en", "transcribe", True) == [50258, 50259, 50360, 50364]
    assert tok.prompt("de", "translate", False) == [50258, 50261, 50359]
    sp2 = bw.special_tokens(bw.dims_for("base"))
    tok2 = Tokenizer(sp2, 51865)
    assert tok2.num_languages == 99 and tok2.prompt("en", "transcribe", True) == [50258, 50259, 50359, 50363]
    with pytest.raises(ValueError):
        tok2.language_token("yue")
