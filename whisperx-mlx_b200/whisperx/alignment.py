"""
Forced alignment for the B200 backend — same public functions as the reference's
whisperx/alignment.py (load_align_model :77, align :113, get_trellis :387, backtrack :447,
backtrack_beam :500, merge_repeats :597), with the numeric core (log-softmax'ed emissions ->
trellis -> backtrack / beam-2) executed by kernel K4 (csrc/wxb_ctc.cu) for ALL segments of a
transcript in one launch, one warp per segment.

Host work that stays in Python: text normalisation, wildcard marking, run-length
merge of the frame path, pandas word / sentence aggregation — written so the resulting dicts
equal the reference's value for value.
"""
import math
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Union

import numpy as np
import pandas as pd
import torch

from .audio import SAMPLE_RATE, load_audio
from .types import AlignedTranscriptionResult, SingleAlignedSegment, SingleSegment, SingleWordSegment
from .utils import LANGUAGES_WITHOUT_SPACES, interpolate_nans

PUNKT_ABBREVIATIONS = ["dr", "vs", "mr", "mrs", "prof"]

# language -> default CTC model (reference tables, alignment.py:31-74; data, not logic)
DEFAULT_ALIGN_MODELS_TORCH = {
    "en": "WAV2VEC2_ASR_BASE_960H", "fr": "VOXPOPULI_ASR_BASE_10K_FR", "de": "VOXPOPULI_ASR_BASE_10K_DE",
    "es": "VOXPOPULI_ASR_BASE_10K_ES", "it": "VOXPOPULI_ASR_BASE_10K_IT",
}
DEFAULT_ALIGN_MODELS_HF = {
    "ja": "jonatasgrosman/wav2vec2-large-xlsr-53-japanese", "zh": "jonatasgrosman/wav2vec2-large-xlsr-53-chinese-zh-cn",
    "nl": "jonatasgrosman/wav2vec2-large-xlsr-53-dutch", "uk": "Yehor/wav2vec2-xls-r-300m-uk-with-small-lm",
    "pt": "jonatasgrosman/wav2vec2-large-xlsr-53-portuguese", "ar": "jonatasgrosman/wav2vec2-large-xlsr-53-arabic",
    "cs": "comodoro/wav2vec2-xls-r-300m-cs-250", "ru": "jonatasgrosman/wav2vec2-large-xlsr-53-russian",
    "pl": "jonatasgrosman/wav2vec2-large-xlsr-53-polish", "hu": "jonatasgrosman/wav2vec2-large-xlsr-53-hungarian",
    "fi": "jonatasgrosman/wav2vec2-large-xlsr-53-finnish", "fa": "jonatasgrosman/wav2vec2-large-xlsr-53-persian",
    "el": "jonatasgrosman/wav2vec2-large-xlsr-53-greek", "tr": "mpoyraz/wav2vec2-xls-r-300m-cv7-turkish",
    "da": "saattrupdan/wav2vec2-xls-r-300m-ftspeech", "he": "imvladikon/wav2vec2-xls-r-300m-hebrew",
    "vi": "nguyenvulebinh/wav2vec2-base-vi", "ko": "kresnik/wav2vec2-large-xlsr-korean",
    "ur": "kingabzpro/wav2vec2-large-xls-r-300m-Urdu", "te": "anuragshas/wav2vec2-large-xlsr-53-telugu",
    "hi": "theainerd/Wav2Vec2-large-xlsr-hindi", "ca": "softcatala/wav2vec2-large-xlsr-catala",
    "ml": "gvs/wav2vec2-large-xlsr-malayalam", "no": "NbAiLab/nb-wav2vec2-1b-bokmaal-v2",
    "nn": "NbAiLab/nb-wav2vec2-1b-nynorsk", "sk": "comodoro/wav2vec2-xls-r-300m-sk-cv8",
    "sl": "anton-l/wav2vec2-large-xlsr-53-slovenian", "hr": "classla/wav2vec2-xls-r-parlaspeech-hr",
    "ro": "gigant/romanian-wav2vec2", "eu": "stefan-it/wav2vec2-large-xlsr-53-basque",
    "gl": "ifrz/wav2vec2-large-xlsr-galician", "ka": "xsway/wav2vec2-large-xlsr-georgian",
    "lv": "jimregan/wav2vec2-large-xlsr-latvian-cv", "tl": "Khalsuu/filipino-wav2vec2-l-xls-r-300m-official",
}


# ------------------------------------------------------------------------------------------------
# model loading (library code: torchaudio / transformers provide the CTC acoustic model)
# ------------------------------------------------------------------------------------------------
def load_align_model(language_code: str, device: str, model_name: Optional[str] = None, model_dir=None,
                     random_init: bool = False, native: bool = True, seed: int = 0):
    """Same contract as alignment.py:77-110: returns (model, {"language","dictionary","type"}).

    For torchaudio bundles of the wav2vec2-BASE architecture (the en / fr / de / es / it defaults) the returned model is the
    B200-native `Wav2Vec2B200` (whisperx/align_model.py): same weights, forward pass on our own kernels, batched over all
    segments by align().  `native=False` keeps the torch module; Hugging Face checkpoints (wav2vec2 LARGE / XLSR, a
    layer-norm feature extractor) stay torch modules.  `random_init=True` builds the bundle's architecture with seeded random
    weights instead of downloading a checkpoint (benchmarks and tests: there is no network)."""
    import torchaudio

    if model_name is None:
        model_name = DEFAULT_ALIGN_MODELS_TORCH.get(language_code) or DEFAULT_ALIGN_MODELS_HF.get(language_code)
        if model_name is None:
            print(f"There is no default alignment model set for this language ({language_code}). "
                  "Please find a wav2vec2.0 model finetuned on this language in https://huggingface.co/models, "
                  "then pass the model name in --align_model [MODEL_NAME]")
            raise ValueError(f"No default align-model for language: {language_code}")

    if model_name in torchaudio.pipelines.__all__:
        kind = "torchaudio"
        bundle = getattr(torchaudio.pipelines, model_name)
        if random_init:
            from .align_model import random_init_torchaudio
            model = random_init_torchaudio(bundle._params, seed)
        else:
            model = bundle.get_model(dl_kwargs={"model_dir": model_dir})
        dictionary = {label.lower(): idx for idx, label in enumerate(bundle.get_labels())}
        params = bundle._params
        base_arch = (params.get("extractor_mode") == "group_norm" and params.get("encoder_embed_dim", 0) <= 1024
                     and not params.get("encoder_layer_norm_first", True))
        if native and base_arch and torch.device(device).type == "cuda":
            from .align_model import Wav2Vec2B200, W2V_BASE_DIMS
            dims = dict(W2V_BASE_DIMS, embed_dim=params["encoder_embed_dim"], n_heads=params["encoder_num_heads"],
                        n_layers=params["encoder_num_layers"], ff_dim=params["encoder_ff_interm_features"],
                        pos_kernel=params["encoder_pos_conv_kernel"], pos_groups=params["encoder_pos_conv_groups"],
                        n_out=params["aux_num_out"])
            model = Wav2Vec2B200(model.state_dict(), device, dims)
        else:
            model = model.to(device)
    else:
        from transformers import Wav2Vec2ForCTC, Wav2Vec2Processor
        try:
            processor = Wav2Vec2Processor.from_pretrained(model_name, cache_dir=model_dir)
            model = Wav2Vec2ForCTC.from_pretrained(model_name, cache_dir=model_dir)
        except Exception as e:
            print(e)
            print("Error loading model from huggingface, check https://huggingface.co/models for finetuned wav2vec2.0 models")
            raise ValueError(f'The chosen align_model "{model_name}" could not be found in huggingface '
                             "(https://huggingface.co/models) or torchaudio (https://pytorch.org/audio/stable/pipelines.html#id14)")
        kind = "huggingface"
        model = model.to(device)
        dictionary = {tok.lower(): idx for tok, idx in processor.tokenizer.get_vocab().items()}
    return model, {"language": language_code, "dictionary": dictionary, "type": kind}


# ------------------------------------------------------------------------------------------------
# kernel-level seams with the reference's signatures
# ------------------------------------------------------------------------------------------------
@dataclass
class Point:
    token_index: int
    time_index: int
    score: float


@dataclass
class Segment:
    label: str
    start: int
    end: int
    score: float

    def __repr__(self):
        return f"{self.label}\t({self.score:4.2f}): [{self.start:5d}, {self.end:5d})"

    @property
    def length(self):
        return self.end - self.start


def _ctx_for(device=None):
    from ._native import get_context
    if device is None:
        idx = torch.cuda.current_device() if torch.cuda.is_available() else 0
    else:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the b200 aligner runs on the GPU only (no CPU fallback)")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return get_context(idx)


def _as_emission(ctx, emission) -> torch.Tensor:
    e = emission if torch.is_tensor(emission) else torch.as_tensor(np.asarray(emission))
    return e.to(ctx.device, dtype=torch.float32).contiguous()


def _tokens_dev(ctx, tokens) -> torch.Tensor:
    return torch.as_tensor(np.asarray(list(tokens), dtype=np.int32)).to(ctx.device)


def get_trellis(emission, tokens, blank_id=0):
    """alignment.py:387-404 on the GPU: returns f32 cuda tensor [T, N]."""
    from ._native import CTC_TRELLIS_ONLY
    ctx = _ctx_for(emission.device if torch.is_tensor(emission) and emission.is_cuda else None)
    e = _as_emission(ctx, emission)
    T, N = e.shape[0], len(tokens)
    res = ctx.ctc_align(e, np.array([0, T]), _tokens_dev(ctx, tokens), np.array([0, N]), blank_id,
                        CTC_TRELLIS_ONLY, want_trellis=True)
    return res["trellis"][:T * N].view(T, N)


def _points_from(res, T: int, t0: int = 0) -> List[Point]:
    tok = res["path_tok"][t0:t0 + T].cpu().numpy()
    lp = res["path_lp"][t0:t0 + T].cpu()
    prob = torch.exp(lp).numpy()  # host exp on the exact log-probs: same rounding as the reference's .exp()
    return [Point(int(tok[t]), t, float(prob[t])) for t in range(T)]


def _run_single(emission, tokens, blank_id, mode):
    ctx = _ctx_for(emission.device if torch.is_tensor(emission) and emission.is_cuda else None)
    e = _as_emission(ctx, emission)
    T, N = e.shape[0], len(tokens)
    res = ctx.ctc_align(e, np.array([0, T]), _tokens_dev(ctx, tokens), np.array([0, N]), blank_id, mode)
    return res, T


def backtrack(trellis, emission, tokens, blank_id=0):
    """alignment.py:447-481.  `trellis` is accepted for signature compatibility; the kernel
    recomputes it from (emission, tokens) on the device, bit-identically."""
    from ._native import CTC_BACKTRACK
    res, T = _run_single(emission, tokens, blank_id, CTC_BACKTRACK)
    if int(res["status"][0]) != 0:
        raise AssertionError("backtrack: ran out of frames before reaching the first token (t > 0 violated)")
    return _points_from(res, T)


def backtrack_beam(trellis, emission, tokens, blank_id=0, beam_width=5):
    """alignment.py:500-579 for beam widths 1..8 (the reference's default is 5; align() calls it with 2, alignment.py:269).
    `trellis` is accepted for signature compatibility; the kernel recomputes it bit-identically."""
    from ._native import ctc_beam_mode
    res, T = _run_single(emission, tokens, blank_id, ctc_beam_mode(beam_width))
    if int(res["status"][0]) != 0:
        return None
    return _points_from(res, T)


def merge_repeats(path: List[Point], transcript: str) -> List[Segment]:
    """Run-length encode the frame path by token index (alignment.py:597-613)."""
    out: List[Segment] = []
    n = len(path)
    lo = 0
    while lo < n:
        hi = lo
        while hi < n and path[hi].token_index == path[lo].token_index:
            hi += 1
        mean_score = sum(path[k].score for k in range(lo, hi)) / (hi - lo)
        out.append(Segment(transcript[path[lo].token_index], path[lo].time_index, path[hi - 1].time_index + 1, mean_score))
        lo = hi
    return out


def merge_words(segments, separator="|"):
    """alignment.py:615-629 (unused by align(); kept for API completeness)."""
    words, lo, hi = [], 0, 0
    while lo < len(segments):
        if hi >= len(segments) or segments[hi].label == separator:
            if lo != hi:
                group = segments[lo:hi]
                total = sum(s.length for s in group)
                words.append(Segment("".join(s.label for s in group), group[0].start, group[-1].end,
                                     sum(s.score * s.length for s in group) / total))
            lo = hi = hi + 1
        else:
            hi += 1
    return words


# ------------------------------------------------------------------------------------------------
# align()
# ------------------------------------------------------------------------------------------------
_PUNKT = None  # (PunktParameters, PunktSentenceTokenizer) once imported, False if nltk is absent


def _sentence_spans(text: str):
    """Punkt sentence spans with the reference's abbreviation list (alignment.py:190-194).  nltk is an
    optional dependency here: without it the whole text is one sentence (documented in DESIGN.md)."""
    global _PUNKT
    if _PUNKT is None:
        try:
            from nltk.tokenize.punkt import PunktParameters, PunktSentenceTokenizer
            _PUNKT = (PunktParameters, PunktSentenceTokenizer)
        except ImportError:
            _PUNKT = False
    if not _PUNKT:
        return [(0, len(text))]
    params = _PUNKT[0]()
    params.abbrev_types = set(PUNKT_ABBREVIATIONS)
    return list(_PUNKT[1](params).span_tokenize(text))


def _prepare_segment(text: str, dictionary: dict, spaced: bool):
    """Character cleaning of one transcript segment (alignment.py:151-201): leading / trailing whitespace is skipped, every
    other character is lower-cased (spaces become "|" in spaced languages) and replaced by the wildcard "*" if the align
    model's dictionary does not know it."""
    lead = len(text) - len(text.lstrip())
    trail = len(text) - len(text.rstrip())
    last_kept = len(text) - trail - 1
    if text.isascii():
        # lower() and the space replacement are 1:1 on ASCII: whole-string operations instead of a per-character loop
        body = text[lead:last_kept + 1].lower()
        if spaced:
            body = body.replace(" ", "|")
        clean_char = [c if c in dictionary else "*" for c in body]
        clean_cdx = list(range(lead, last_kept + 1))
    else:
        clean_char, clean_cdx = [], []
        for cdx, ch in enumerate(text):
            if cdx < lead or cdx > last_kept:
                continue
            c = ch.lower()
            if spaced:
                c = c.replace(" ", "|")
            clean_char.append(c if c in dictionary else "*")
            clean_cdx.append(cdx)
    words = text.split(" ") if spaced else text
    return {"clean_char": clean_char, "clean_cdx": clean_cdx, "clean_wdx": list(range(len(words))),
            "sentence_spans": _sentence_spans(text)}


def _emissions_for(model, model_type: str, wave: torch.Tensor, lengths, device):
    with torch.inference_mode():
        if model_type == "torchaudio":
            logits, _ = model(wave.to(device), lengths=lengths)
        elif model_type == "huggingface":
            logits = model(wave.to(device)).logits
        else:
            raise NotImplementedError(f"Align model of type {model_type} not supported.")
    return logits[0]


def _pinned(ctx, name: str, n: int, dtype) -> torch.Tensor:
    """Pinned host buffers kept on the Context between align() calls (cudaHostAlloc per call costs more than the copies)."""
    cache = ctx.__dict__.setdefault("_align_pinned", {})
    t = cache.get(name)
    if t is None or t.numel() < n or t.dtype != dtype:
        t = torch.empty(max(n, 1024), dtype=dtype).pin_memory()
        cache[name] = t
    return t


# align() runs its segments as a pipeline of at most this many groups of at least this many segments (GPU: forward + K4 of one
# group while the host assembles the previous one)
PIPELINE_GROUPS = 2     # measured on 60 x 30 s segments (tools/e2e_profile.py, E2E_GROUPS): 1 group 85.7 ms, 2: 78.3, 3: 81.3, 4: 87.6, 6: 99.1
PIPELINE_MIN_SEGMENTS = 15  # per group: below this one batched forward is the better deal
TRACE = None  # tools/e2e_profile.py sets this to a list: align() then appends (label, perf_counter()) at its pipeline points


def _mark(label: str):
    if TRACE is not None:
        import time
        TRACE.append((label, time.perf_counter()))


def align(
    transcript: Iterable[SingleSegment],
    model: torch.nn.Module,
    align_model_metadata: dict,
    audio: Union[str, np.ndarray, torch.Tensor],
    device: str,
    interpolate_method: str = "nearest",
    return_char_alignments: bool = False,
    print_progress: bool = False,
    combined_progress: bool = False,
) -> AlignedTranscriptionResult:
    """Drop-in for alignment.py:113-380.  Emissions of every alignable segment are gathered on the
    GPU, log-softmax'ed and aligned by ONE K4 launch (beam-2, as the reference), then the host
    builds the same char / word / sentence dicts."""
    from ._native import CTC_BEAM2

    _mark("start")
    if not torch.is_tensor(audio):
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = torch.from_numpy(audio)
    if audio.dim() == 1:
        audio = audio.unsqueeze(0)
    max_duration = audio.shape[1] / SAMPLE_RATE

    dictionary = align_model_metadata["dictionary"]
    lang = align_model_metadata["language"]
    model_type = align_model_metadata["type"]
    spaced = lang not in LANGUAGES_WITHOUT_SPACES
    ctx = _ctx_for(device)

    transcript = list(transcript)
    total = len(transcript)
    prepared = []
    for sdx, seg in enumerate(transcript):
        if print_progress:
            base = ((sdx + 1) / total) * 100
            print(f"Progress: {(50 + base / 2) if combined_progress else base:.2f}%...")
        prepared.append(_prepare_segment(seg["text"], dictionary, spaced))

    blank_id = 0
    for ch, code in dictionary.items():
        if ch == "[pad]" or ch == "<pad>":
            blank_id = code

    _mark("segments prepared")
    # ---- pass 1: emissions of every alignable segment, kept on the device ------------------
    ascii_lut = np.full(128, -1, dtype=np.int32)  # dictionary code of every single ASCII character (-1 = wildcard)
    for ch, code in dictionary.items():
        if len(ch) == 1 and ord(ch) < 128:
            ascii_lut[ord(ch)] = code
    jobs = []  # (sdx, text_clean, tokens, T)
    emis_parts, tok_parts = [], []
    skip_reason = {}
    native = bool(getattr(model, "is_b200_native", False))  # Wav2Vec2B200: ONE batched forward over all segments
    native_waves = []
    for sdx, seg in enumerate(transcript):
        t1, t2 = seg["start"], seg["end"]
        if len(prepared[sdx]["clean_char"]) == 0:
            skip_reason[sdx] = "no characters in this segment found in model dictionary, resorting to original..."
            continue
        if t1 >= max_duration:
            skip_reason[sdx] = "original start time longer than audio duration, skipping..."
            continue
        text_clean = "".join(prepared[sdx]["clean_char"])
        if text_clean.isascii():  # one table look-up per segment instead of a dict look-up per character
            tokens = ascii_lut[np.frombuffer(text_clean.encode("ascii"), dtype=np.uint8)]
        else:
            tokens = [dictionary.get(c, -1) for c in text_clean]
        f1, f2 = int(t1 * SAMPLE_RATE), int(t2 * SAMPLE_RATE)
        wave = audio[:, f1:f2]
        if native:
            if wave.shape[0] != 1:
                raise ValueError("the B200 alignment model takes mono audio")
            from .align_model import frames_for, MIN_SAMPLES
            native_waves.append(wave[0].numpy() if not wave.is_cuda else wave[0].cpu().numpy())
            tok_parts.append(np.asarray(tokens, dtype=np.int32))
            jobs.append((sdx, text_clean, tokens, frames_for(max(wave.shape[-1], MIN_SAMPLES)), 1))
            continue
        lengths = None
        if wave.shape[-1] < 400:  # minimum wav2vec2 input (alignment.py:243-249)
            lengths = torch.as_tensor([wave.shape[-1]]).to(device)
            wave = torch.nn.functional.pad(wave, (0, 400 - wave.shape[-1]))
        logits = _emissions_for(model, model_type, wave, lengths, device)
        logits = logits.to(ctx.device, dtype=torch.float32)
        emis_parts.append(logits)
        tok_parts.append(np.asarray(tokens, dtype=np.int32))
        jobs.append((sdx, text_clean, tokens, logits.shape[0], wave.size(0)))

    # ---- pass 2: K4 on the GPU, host assembly of the dicts ------------------------------------
    # The jobs run as a short pipeline of groups: while the GPU does the wav2vec2 forward + log-softmax + K4 of group i + 1, the
    # host assembles the char / word / sentence dicts of group i (path read-back through pinned buffers and an event, no
    # blocking copy in between).  One group when there are few segments or a torch model produced the emissions.
    assembled = {}   # sdx -> list of aligned segments
    failed = {}      # sdx -> message
    n_groups = min(PIPELINE_GROUPS, max(1, len(jobs) // PIPELINE_MIN_SEGMENTS)) if native else 1
    if n_groups == 2:
        # two groups, the first twice the size of the second: the host work left over when the GPU is done is the assembly of
        # the LAST group only, and a small batch costs the forward little once it is a third of the job
        bounds = [0, (2 * len(jobs) + 2) // 3, len(jobs)]
    else:
        bounds = [len(jobs) * g // n_groups for g in range(n_groups + 1)]
    d2h_bytes = 0
    h2d_bytes = 0
    stats_sum = {"flops": 0.0, "frames": 0, "segments": 0}

    def launch(g, a, b):
        nonlocal h2d_bytes
        _mark("launch %d: begin" % g)
        toks = np.concatenate(tok_parts[a:b]) if b > a else np.zeros(0, np.int32)
        t_off = np.concatenate([[0], np.cumsum([j[3] for j in jobs[a:b]])]).astype(np.int32)
        n_off = np.concatenate([[0], np.cumsum([len(t) for t in tok_parts[a:b]])]).astype(np.int32)
        sum_t = int(t_off[-1])
        pin = _pinned(ctx, "tok%d" % (g % 2), max(len(toks), 1), torch.int32)
        pin[: len(toks)].copy_(torch.from_numpy(toks))
        tok_dev = torch.empty(max(len(toks), 1), dtype=torch.int32, device=ctx.device)
        tok_dev[: len(toks)].copy_(pin[: len(toks)], non_blocking=True)  # before the forward: nothing below waits for the stream
        if native:
            emis, _ = model.emissions(native_waves[a:b])  # pinned upload -> batched wav2vec2 forward (csrc/wxb_w2v.cu) -> [sum T, V]
            h2d_bytes += int(model.last_stats.get("h2d_bytes", 0)) + int(toks.nbytes)
            for k2 in stats_sum:
                stats_sum[k2] += model.last_stats.get(k2, 0)
        else:
            emis = torch.cat(emis_parts[a:b], 0).contiguous()
        ctx.log_softmax_rows_(emis)  # alignment.py:258
        res = ctx.ctc_align(emis, t_off, tok_dev[: len(toks)], n_off, blank_id, CTC_BEAM2)
        slot = g % 2  # two sets of read-back buffers: group g + 1 is launched before group g is read
        h_status = _pinned(ctx, "status%d" % slot, b - a, torch.int32)
        h_tok = _pinned(ctx, "ptok%d" % slot, max(sum_t, 1), torch.int32)
        h_lp = _pinned(ctx, "plp%d" % slot, max(sum_t, 1), torch.float32)
        h_status[: b - a].copy_(res["status"], non_blocking=True)
        h_tok[:sum_t].copy_(res["path_tok"], non_blocking=True)
        h_lp[:sum_t].copy_(res["path_lp"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _mark("launch %d: enqueued" % g)
        return (a, b, t_off, sum_t, h_status, h_tok, h_lp, ev, (emis, tok_dev, res))

    def finish(h):
        nonlocal d2h_bytes
        a, b, t_off, sum_t, h_status, h_tok, h_lp, ev, _keep = h
        _mark("finish: wait for the GPU")
        ev.synchronize()
        _mark("finish: paths on the host")
        status = h_status[: b - a].numpy()
        path_tok = h_tok[:sum_t].numpy()
        path_prob = torch.exp(h_lp[:sum_t]).numpy()
        d2h_bytes += int(status.nbytes + path_tok.nbytes + path_prob.nbytes)
        for k in range(a, b):
            sdx, text_clean, _tokens, T, n_channels = jobs[k]
            seg = transcript[sdx]
            t1, t2, text = seg["start"], seg["end"], seg["text"]
            if int(status[k - a]) != 0:
                failed[sdx] = "backtrack failed, resorting to original..."
                continue
            f0, f1 = int(t_off[k - a]), int(t_off[k - a + 1])
            runs = _merge_runs(path_tok[f0:f1], path_prob[f0:f1])
            ratio = (t2 - t1) * n_channels / (T - 1)
            assembled[sdx] = _assemble(text, prepared[sdx], runs, ratio, t1, spaced, interpolate_method, return_char_alignments)

    _mark("tokens / waves listed")
    pending = None
    for g in range(n_groups):
        if bounds[g + 1] == bounds[g]:
            continue
        if native and model.ctx is not ctx:
            raise RuntimeError("the alignment model lives on another GPU than `device`")
        h = launch(g, bounds[g], bounds[g + 1])
        if pending is not None:
            finish(pending)
        pending = h
    if pending is not None:
        finish(pending)
    _mark("all groups assembled")
    if native and jobs:  # totals of this align() call (the groups' forwards each wrote their own)
        model.last_stats.update(d2h_bytes=d2h_bytes, h2d_bytes=h2d_bytes, **stats_sum)

    aligned_segments: List[SingleAlignedSegment] = []
    for sdx, seg in enumerate(transcript):
        t1, t2, text = seg["start"], seg["end"], seg["text"]
        if sdx in assembled:
            aligned_segments += assembled[sdx]
            continue
        plain: SingleAlignedSegment = {"start": t1, "end": t2, "text": text, "words": [], "chars": None}
        if return_char_alignments:
            plain["chars"] = []
        print(f'Failed to align segment ("{text}"): {skip_reason.get(sdx) or failed.get(sdx)}')
        aligned_segments.append(plain)

    word_segments: List[SingleWordSegment] = []
    for seg in aligned_segments:
        word_segments += seg["words"]
    _mark("end")
    return {"segments": aligned_segments, "word_segments": word_segments}


def _assemble(text, prep, runs, ratio, t1, spaced, interpolate_method, return_char_alignments):
    """Char -> word -> sentence aggregation (alignment.py:281-373).  A segment that is ONE sentence (always the case when
    nltk is absent, and for most ASR segments) takes the numpy path below, which reproduces the pandas results value for
    value (tests/test_host_cpu.py compares the two on random inputs); everything else goes through the reference's own
    pandas primitives."""
    spans = prep["sentence_spans"]
    if runs and not isinstance(runs, tuple):  # a list of Segment objects (merge_repeats output)
        runs = ([g.start for g in runs], [g.end for g in runs], [g.score for g in runs])
    if len(spans) == 1:
        return _assemble_single_sentence(text, prep, runs, ratio, t1, spaced, return_char_alignments)
    runs = tuple(np.asarray(r).tolist() for r in runs)  # plain Python ints / floats, as merge_repeats produces them
    char_segments = [Segment("", a, b, v) for a, b, v in zip(*runs)]  # the pandas path only reads start / end / score
    return _assemble_pandas(text, prep, char_segments, ratio, t1, spaced, interpolate_method, return_char_alignments)


def _round3(a: np.ndarray) -> np.ndarray:
    """Python's round(x, 3) (correctly rounded decimal, what the reference applies to every character time) for an array:
    numpy's multiply-rint-divide gives the same double unless x sits on a rounding boundary; only those go through round()."""
    out = np.round(a, 3)
    y = a * 1000.0
    for i in np.flatnonzero(np.abs(y - np.floor(y) - 0.5) < 1e-6).tolist():
        out[i] = round(float(a[i]), 3)
    return out


def _assemble_single_sentence(text, prep, runs, ratio, t1, spaced, return_char_alignments):
    """The pandas pipeline of alignment.py:308-373 for a single sentence span on numpy arrays, a handful of array operations
    per SEGMENT instead of several pandas calls per WORD: per word fmin(start) / fmax(end) / round(nanmean(score), 3) through
    ufunc.reduceat over the word's non-space characters (the mean exactly as pandas' nanmean: NaNs count as 0 in one sum,
    divided by the number of non-NaN values), then the one-row groupby(start, end): a row whose start or end is NaN is
    dropped, otherwise the row comes back with native Python floats.  `runs` = (first frame, end frame, mean prob) lists of
    the merged path, one entry per clean character."""
    n = len(text)
    r_start, r_end, r_score = runs
    start = np.full(n, np.nan)
    end = np.full(n, np.nan)
    score = np.full(n, np.nan)
    cdx = prep["clean_cdx"]
    if len(cdx):
        start[cdx] = _round3(np.asarray(r_start, dtype=np.float64) * ratio + t1)
        end[cdx] = _round3(np.asarray(r_end, dtype=np.float64) * ratio + t1)
        score[cdx] = _round3(np.asarray(r_score, dtype=np.float64))
    s0, s1 = prep["sentence_spans"][0]
    lo, hi = max(int(s0), 0), min(int(s1), n - 1)  # `.loc` selection: both ends inclusive, clipped to the rows that exist
    if hi < lo:
        return []
    is_space = np.frombuffer(text.encode("utf-32-le"), dtype=np.uint32) == 32
    if spaced:
        widx = np.cumsum(is_space) - int(is_space[0])  # a new word starts AT each " " after the first character
    else:
        widx = np.arange(n)
    sel = slice(lo, hi + 1)
    words = []
    sp = is_space[sel]
    wsel = widx[sel]
    first = np.concatenate([[0], np.flatnonzero(wsel[1:] != wsel[:-1]) + 1])  # reduceat offsets: one group per word of the selection
    st_m = np.where(sp, np.nan, start[sel])   # spaces take no part in a word's times / score
    en_m = np.where(sp, np.nan, end[sel])
    sc_m = np.where(sp, np.nan, score[sel])
    w_start = np.fmin.reduceat(st_m, first)
    w_end = np.fmax.reduceat(en_m, first)
    n_ns = np.add.reduceat((~sp).astype(np.int64), first)   # non-space characters per word
    # score: pandas' nanmean = ONE ndarray.sum over the word's non-space characters with NaNs counted as 0, divided by the
    # number of non-NaN values, then round(., 3).  The sum must associate exactly as ndarray.sum does, or a mean that sits on a
    # rounding boundary of round(., 3) (every other two-letter word: scores are multiples of 0.001) rounds the other way:
    #   * fewer than 8 addends: ndarray.sum adds left to right.  All such words are summed at once, column by column over a
    #     [words, 7] gather of the compacted scores (trailing zeros are exact);
    #   * 8 or more: ndarray.sum is pairwise with its own blocking.  reduceat's order differs from it by an ulp at most, which
    #     matters only on a rounding boundary: those words (and only those) are re-summed with ndarray.sum itself.
    sc_ns = score[sel][~sp]
    ok_ns = ~np.isnan(sc_ns)
    z_ns = np.where(ok_ns, sc_ns, 0.0)
    j1 = np.cumsum(n_ns)
    j0 = j1 - n_ns                                             # word k owns z_ns[j0[k]:j1[k]]
    ok_cum = np.concatenate([[0], np.cumsum(ok_ns)])
    cnt = ok_cum[j1] - ok_cum[j0]                              # non-NaN scores per word
    zp = np.concatenate([z_ns, np.zeros(7)])
    col = np.arange(7)
    vals = np.where(col[None, :] < n_ns[:, None], zp[j0[:, None] + col[None, :]], 0.0)
    acc = vals[:, 0].copy()
    for c in range(1, 7):
        acc += vals[:, c]
    with np.errstate(invalid="ignore", divide="ignore"):
        w_score = np.round(acc / cnt, 3)
        big = np.flatnonzero((n_ns >= 8) & (cnt > 0))
        if big.size:
            ok_m = ~np.isnan(sc_m)
            mean_fast = np.add.reduceat(np.where(ok_m, sc_m, 0.0), first)[big] / cnt[big]
            y = mean_fast * 1000.0
            w_score[big] = np.round(mean_fast, 3)
            for k in big[np.abs(y - np.floor(y) - 0.5) < 1e-6].tolist():
                w_score[k] = np.round(z_ns[j0[k]:j1[k]].sum() / cnt[k], 3)
    wtexts = None
    if spaced:
        # a group = a space and the characters up to the next space: the pieces of str.split(" "), minus the empty piece in
        # front of a leading space
        pieces = text[lo:hi + 1].split(" ")
        if sp[0]:
            del pieces[0]
        if len(pieces) == len(first):
            wtexts = [p.strip() for p in pieces]
    if wtexts is None:
        bounds = (first + lo).tolist() + [hi + 1]
        wtexts = [text[c:e].strip() for c, e in zip(bounds[:-1], bounds[1:])]
    full = (~np.isnan(w_start)) & (~np.isnan(w_end)) & (cnt > 0)
    if full.all():  # the usual case: every word has its times and a score
        words = [{"word": t, "start": a, "end": b, "score": c} for t, a, b, c in zip(wtexts, w_start, w_end, w_score) if t]
    else:
        has_s, has_e, has_c = (~np.isnan(w_start)).tolist(), (~np.isnan(w_end)).tolist(), (cnt > 0).tolist()
        for k, wtext in enumerate(wtexts):
            if not wtext:
                continue
            entry = {"word": wtext}
            if has_s[k]:
                entry["start"] = w_start[k]
            if has_e[k]:
                entry["end"] = w_end[k]
            if has_c[k]:
                entry["score"] = w_score[k]
            words.append(entry)
    ns = np.flatnonzero(~sp) + lo
    sent_start = np.fmin.reduce(start[sel]) if hi >= lo else np.nan
    sent_end = np.fmax.reduce(end[ns]) if ns.size else np.nan
    if np.isnan(sent_start) or np.isnan(sent_end):
        return []  # groupby drops rows with a NaN key (a single row has no neighbour to interpolate from)
    out = {"start": float(sent_start), "end": float(sent_end), "text": text[s0:s1], "words": words}
    if return_char_alignments:
        chars = []
        st_l, en_l, sc_l = start.tolist(), end.tolist(), score.tolist()
        for c in range(lo, hi + 1):
            rec = {"char": text[c]}
            if st_l[c] == st_l[c]:
                rec["start"] = st_l[c]
            if en_l[c] == en_l[c]:
                rec["end"] = en_l[c]
            if sc_l[c] == sc_l[c]:
                rec["score"] = sc_l[c]
            chars.append(rec)
        out["chars"] = chars
    return [out]


def _merge_runs(path_tok: np.ndarray, path_prob: np.ndarray):
    """merge_repeats (alignment.py:597-613) straight from the kernel's per-frame arrays: (first frame, end frame, mean prob)
    of every run of equal token index, as three lists.  The mean is the left-to-right double sum of the per-frame float
    probabilities over the run length, exactly what merging T Point objects gives."""
    T = len(path_tok)
    if T == 0:
        return [], [], []
    cut = np.flatnonzero(path_tok[1:] != path_tok[:-1]) + 1
    first = np.concatenate([[0], cut])
    hi_arr = np.concatenate([cut, [T]])
    length = hi_arr - first
    # The probabilities are float32 values: a double sum of a few of them is EXACT (no rounding at all, hence independent of
    # the summation order and of Python's compensated sum) whenever the binary exponents inside the run span less than
    # 53 - 24 - log2(length) bits.  Those runs are summed with one reduceat; the rest (a run mixing ~1 and ~1e-9) take
    # Python's own sum(), which is what the reference's merge_repeats evaluates.
    p64 = path_prob.astype(np.float64)
    _, ex = np.frexp(p64)
    ex = np.where(p64 == 0.0, 10000, ex)  # zeros constrain nothing
    e_min = np.minimum.reduceat(ex, first)
    e_max = np.maximum.reduceat(np.where(p64 == 0.0, -10000, ex), first)
    exact = (e_max - e_min <= 24) & (length <= 16)
    sums = np.add.reduceat(p64, first)
    mean = sums / length
    if not exact.all():
        prob = p64.tolist()
        lo, hi = first.tolist(), hi_arr.tolist()
        for k in np.flatnonzero(~exact).tolist():
            mean[k] = sum(prob[lo[k]:hi[k]]) / (hi[k] - lo[k])
    return first, hi_arr, mean


def _merge_repeats_arrays(path_tok: np.ndarray, path_prob: np.ndarray, transcript: str) -> List[Segment]:
    lo, hi, sc = (np.asarray(r).tolist() for r in _merge_runs(path_tok, path_prob))
    tok = path_tok.tolist()
    return [Segment(transcript[tok[a]], a, b, v) for a, b, v in zip(lo, hi, sc)]


def _assemble_pandas(text, prep, char_segments, ratio, t1, spaced, interpolate_method, return_char_alignments):
    """Char -> word -> sentence aggregation (alignment.py:281-373), same pandas primitives so the
    float results (min / max / mean, NaN handling, groupby ordering) are the reference's."""
    pos_of = {cdx: k for k, cdx in enumerate(prep["clean_cdx"])}
    rows = []
    widx = 0
    for cdx, ch in enumerate(text):
        start = end = score = None
        k = pos_of.get(cdx)
        if k is not None:
            cs = char_segments[k]
            start = round(cs.start * ratio + t1, 3)
            end = round(cs.end * ratio + t1, 3)
            score = round(cs.score, 3)
        rows.append({"char": ch, "start": start, "end": end, "score": score, "word-idx": widx})
        if not spaced:
            widx += 1
        elif cdx == len(text) - 1 or text[cdx + 1] == " ":
            widx += 1
    chars = pd.DataFrame(rows)
    chars["sentence-idx"] = None

    sentences = []
    for sidx, (s0, s1) in enumerate(prep["sentence_spans"]):
        sel = (chars.index >= s0) & (chars.index <= s1)
        cur = chars.loc[sel]
        chars.loc[sel, "sentence-idx"] = sidx
        non_space = cur[cur["char"] != " "]
        words = []
        for w in cur["word-idx"].unique():
            wc = cur.loc[cur["word-idx"] == w]
            wtext = "".join(wc["char"].tolist()).strip()
            if len(wtext) == 0:
                continue
            wc = wc[wc["char"] != " "]
            w_start, w_end = wc["start"].min(), wc["end"].max()
            w_score = round(wc["score"].mean(), 3)
            entry = {"word": wtext}
            if not np.isnan(w_start):
                entry["start"] = w_start
            if not np.isnan(w_end):
                entry["end"] = w_end
            if not np.isnan(w_score):
                entry["score"] = w_score
            words.append(entry)
        sent = {"text": text[s0:s1], "start": cur["start"].min(), "end": non_space["end"].max(), "words": words}
        if return_char_alignments:
            cc = cur[["char", "start", "end", "score"]].copy()
            cc.fillna(-1, inplace=True)
            sent["chars"] = [{k: v for k, v in rec.items() if v != -1} for rec in cc.to_dict("records")]
        sentences.append(sent)

    df = pd.DataFrame(sentences)
    df["start"] = interpolate_nans(df["start"], method=interpolate_method)
    df["end"] = interpolate_nans(df["end"], method=interpolate_method)
    agg = {"text": " ".join if spaced else "".join, "words": "sum"}
    if return_char_alignments:
        agg["chars"] = "sum"
    df = df.groupby(["start", "end"], as_index=False).agg(agg)
    return df.to_dict("records")


_PINNED: Dict[int, Any] = {}


def _pinned_rows(device_index: int, rows: int, V: int) -> torch.Tensor:
    """[rows, V] f32 view of a grow-only pinned staging buffer; waits for the copy that read it last."""
    ent = _PINNED.get(device_index)
    if ent is None or ent[0].numel() < rows * V:
        ent = [torch.empty(max(rows * V, 1), dtype=torch.float32).pin_memory(), None]
        _PINNED[device_index] = ent
    if ent[1] is not None:
        ent[1].synchronize()
    return ent[0][: rows * V].view(rows, V)


def align_from_emissions(emission_logits: List[np.ndarray], token_lists: List[List[int]], blank_id: int = 0,
                         device_index: int = 0, beam: bool = True):
    """Numeric core of align() for callers that already hold the CTC model's output (host arrays):
    pinned H2D copy -> log_softmax (alignment.py:258) -> K4 trellis + beam-2 / backtrack for all
    segments in one launch -> D2H.  Returns per segment (status, token_index[T], prob[T])."""
    from ._native import CTC_BACKTRACK, CTC_BEAM2, get_context
    ctx = get_context(device_index)
    T = [int(e.shape[0]) for e in emission_logits]
    V = int(emission_logits[0].shape[1])
    host = _pinned_rows(device_index, sum(T), V)  # staging kept between calls: page-locking is the expensive part
    off = 0
    for e in emission_logits:
        host[off:off + e.shape[0]] = torch.from_numpy(np.ascontiguousarray(e, dtype=np.float32))
        off += e.shape[0]
    emis = host.to(ctx.device, non_blocking=True)
    _PINNED[device_index][1] = torch.cuda.Event()
    _PINNED[device_index][1].record()
    ctx.log_softmax_rows_(emis)
    t_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int32)
    n_off = np.concatenate([[0], np.cumsum([len(t) for t in token_lists])]).astype(np.int32)
    tok = torch.from_numpy(np.concatenate([np.asarray(t, dtype=np.int32) for t in token_lists])).pin_memory()
    res = ctx.ctc_align(emis, t_off, tok.to(ctx.device, non_blocking=True), n_off, blank_id, CTC_BEAM2 if beam else CTC_BACKTRACK)
    status = res["status"].cpu().numpy()
    ptok = res["path_tok"].cpu().numpy()
    prob = torch.exp(res["path_lp"].cpu()).numpy()
    return [(int(status[k]), ptok[t_off[k]:t_off[k + 1]], prob[t_off[k]:t_off[k + 1]]) for k in range(len(T))]
