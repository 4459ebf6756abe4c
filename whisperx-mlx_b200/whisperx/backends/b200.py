"""
B200WhisperBackend — the new entry in whisperx/backends: log-mel -> encoder -> batched greedy
KV-cache decode on hand-written sm_100a kernels (libwxb200.so), behind the reference's backend
interface (whisperx/backends/base.py:8-57) and the duck-typed `transcribe_batch` fast path
(whisperx/asr.py:67-87; shape per whisperx/backends/mlx_lightning.py:82-119).

Segment convention: upstream-WhisperX style — `without_timestamps` prompt, one
segment per VAD chunk: {"text", "start": round(chunk.start, 3), "end": round(chunk.end, 3)} plus
"tokens", "avg_logprob", "no_speech_prob".  Empty text drops the segment (mlx_lightning.py:199-200).
"""
import os
import warnings
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch

from ..audio import N_SAMPLES, SAMPLE_RATE, load_audio, mel_filters
from ..tokenizer import LANGUAGE_CODES, Tokenizer
from ..types import TranscriptionResult
from . import b200_weights as bw
from .base import WhisperBackend


MAX_BATCH = 64  # sequences per wxb_decode_greedy call


class B200WhisperBackend(WhisperBackend):
    def __init__(self, model: str, device: str = "cuda", device_index: int = 0, compute_type: str = "bfloat16",
                 download_root: Optional[str] = None, local_files_only: bool = False, threads: int = 4,
                 asr_options: Optional[dict] = None, language: Optional[str] = None, task: str = "transcribe",
                 weights: Optional[Union[str, Dict[str, torch.Tensor]]] = None, seed: int = 0,
                 tokenizer: Optional[Tokenizer] = None, **kwargs):
        from .._native import get_context
        if str(device).startswith("cpu"):
            raise RuntimeError("the b200 backend runs on CUDA sm_100a only; there is no CPU fallback")
        self.model_name = bw.canonical_name(model)
        self.dims = bw.dims_for(self.model_name)
        self.ctx = get_context(device_index)
        self.device = self.ctx.device
        self.compute_type = "bfloat16"  # kernels compute in bf16 with fp32 accumulation whatever was asked
        self.language = language
        self.task = task or "transcribe"
        self.options = dict(suppress_blank=True, suppress_tokens=[], sample_len=self.dims["n_text_ctx"] // 2,
                            without_timestamps=True, dtw_word_timestamps=False, alignment_heads=None)
        if asr_options:
            for k in ("suppress_blank", "suppress_tokens", "sample_len", "without_timestamps", "dtw_word_timestamps", "alignment_heads"):
                if k in asr_options and asr_options[k] is not None:
                    self.options[k] = asr_options[k]
        self.options["without_timestamps"] = bool(self.options["without_timestamps"])
        self.specials = bw.special_tokens(self.dims)
        vocab = kwargs.get("vocab")  # path of multilingual.tiktoken / vocab.json
        self.tokenizer = tokenizer or (Tokenizer.from_file(vocab, self.specials, self.dims["n_vocab"]) if vocab
                                       else Tokenizer(self.specials, self.dims["n_vocab"]))
        if -1 in list(self.options["suppress_tokens"]):
            # "-1" = the tokenizer's non-speech symbol set + the special tokens that must never be sampled (the reference's default)
            rest = [t for t in self.options["suppress_tokens"] if t >= 0]
            if self.tokenizer.ranks is not None:
                sp = self.specials
                # OpenAI decoding.py _get_suppress_tokens: + transcribe, translate, sot, startoflm, startofprev, nospeech
                extra = [sp["transcribe"], sp["translate"], sp["sot"], sp["transcribe"] + 1, sp["transcribe"] + 2, sp["no_speech"]]
                self.options["suppress_tokens"] = sorted(set(rest) | set(self.tokenizer.non_speech_tokens()) | set(extra))
            else:
                warnings.warn("suppress_tokens=[-1] (the tokenizer's non-speech set) needs a real vocabulary (vocab=<multilingual.tiktoken | "
                              "vocab.json>); no symbol token is suppressed")
                self.options["suppress_tokens"] = rest
        if self.tokenizer.ranks is not None:
            self.specials["blank"] = self.tokenizer.encode(" ")[0]  # SuppressBlank: encode(" ")
        if weights is None and download_root and os.path.isdir(os.path.join(download_root, model)):
            weights = os.path.join(download_root, model)  # a local copy of the checkpoint, as `download_root` holds in the reference
        self.kernel_weights = self._load_weights(weights, seed)
        bw.validate_kernel_weights(self.kernel_weights, self.dims)
        self._bind()
        self._filters = mel_filters(self.device, self.dims["n_mels"])
        self.align_model = kwargs.get("align_model")  # optional (model, metadata) for _align_words
        self.last_stats: Dict[str, Any] = {}

    # ------------------------------------------------------------------ weights
    def _load_weights(self, weights, seed):
        if weights is None:
            warnings.warn(f"no checkpoint given for '{self.model_name}': using seeded random-init weights "
                          "(pass weights=<state dict or path> for a trained model)")
            return bw.random_kernel_weights_on_device(self.dims, self.device, seed=seed)
        if isinstance(weights, (str, os.PathLike)):
            weights = bw.load_checkpoint(os.fspath(weights))  # HF safetensors dir / file, OpenAI .pt, mlx-community weights.npz
        if any(k.startswith("model.encoder.") for k in weights):
            weights = bw.from_hf_state_dict(weights)
        if "encoder.conv1.weight" in weights:
            got = bw.infer_dims(weights)
            if got != self.dims:
                diff = {k: (got[k], self.dims[k]) for k in got if got[k] != self.dims[k]}
                raise ValueError(f"checkpoint does not match the '{self.model_name}' architecture: (checkpoint, expected) = {diff}")
        if "enc.conv1.w" in weights:  # already in kernel layout: cast to the dtypes the kernels read
            exp = bw.expected_kernel_tensors(self.dims)
            return {k: v.to(device=self.device, dtype=exp[k][1] if k in exp else v.dtype).contiguous() for k, v in weights.items()}
        return bw.to_kernel_layout(weights, self.dims, self.device)

    def _bind(self):
        """Make this backend's model the one resident in the (process-wide, per-GPU) context.  Several backends may share a
        GPU: whichever runs re-binds its own weight table first (a pointer table, no copies), so one backend can never
        transcribe with another one's weights or dims."""
        if self.ctx.model_owner is not self:
            self.ctx.set_model(self.dims, self.kernel_weights, owner=self)

    # ------------------------------------------------------------------ interface properties
    @property
    def supported_languages(self) -> List[str]:
        return list(LANGUAGE_CODES[: self.tokenizer.num_languages]) if self.is_multilingual else ["en"]

    @property
    def is_multilingual(self) -> bool:
        return self.dims["n_vocab"] >= 51865

    # ------------------------------------------------------------------ device-resident hot path
    def upload_chunks(self, chunks: List[np.ndarray]):
        """Pinned staging + one H2D copy.  Returns (audio_dev, offsets, lengths)."""
        lens = np.array([min(len(c), N_SAMPLES) for c in chunks], dtype=np.int32)
        offs = np.zeros(len(chunks), dtype=np.int64)
        if len(chunks) > 1:
            offs[1:] = np.cumsum(lens[:-1])
        total = int(lens.sum())
        # the pinned staging buffer is kept between calls (page-locking 100+ MB costs tens of ms); the copy that
        # read it last has completed before it is overwritten
        if getattr(self, "_staging", None) is None or self._staging.numel() < max(total, 1):
            self._staging = torch.empty(max(total, 1), dtype=torch.float32).pin_memory()
            self._staging_done = None
        if self._staging_done is not None:
            self._staging_done.synchronize()
        hv = self._staging.numpy()
        dev = torch.empty(max(total, 1), dtype=torch.float32, device=self.device)
        # Groups of chunks: host copy into the pinned buffer, then an asynchronous H2D of that range, so the PCIe
        # transfer of one group runs under the host copy of the next.  Chunks that are back-to-back views of one
        # float32 array (what transcribe() cuts) are copied as one range with torch's multi-threaded copy.
        n, group = len(chunks), 12
        for g0 in range(0, n, group):
            g1 = min(n, g0 + group)
            lo, hi = int(offs[g0]), int(offs[g1 - 1] + lens[g1 - 1])
            if hi > lo:
                run = self._contiguous_run(chunks[g0:g1], lens[g0:g1])
                if run is not None:
                    self._staging[lo:hi].copy_(torch.from_numpy(run))
                else:
                    for c, o, l in zip(chunks[g0:g1], offs[g0:g1], lens[g0:g1]):
                        hv[o:o + l] = np.asarray(c[:l], dtype=np.float32)
                dev[lo:hi].copy_(self._staging[lo:hi], non_blocking=True)
        self._staging_done = torch.cuda.Event()
        self._staging_done.record()
        return dev, offs, lens

    @staticmethod
    def _contiguous_run(chunks, lens):
        """The single float32 array the (clipped) chunks tile back to back in memory, or None."""
        first = chunks[0]
        if not (isinstance(first, np.ndarray) and first.dtype == np.float32 and first.ndim == 1 and first.flags.c_contiguous):
            return None
        addr = first.__array_interface__["data"][0]
        start = addr
        for c, l in zip(chunks, lens):
            if not (isinstance(c, np.ndarray) and c.dtype == np.float32 and c.ndim == 1 and c.flags.c_contiguous):
                return None
            if c.__array_interface__["data"][0] != addr or len(c) != int(l):
                return None  # a gap, a reordering, or a chunk clipped to 30 s
            addr += int(l) * 4
        total = (addr - start) // 4
        return np.lib.stride_tricks.as_strided(first, shape=(total,), strides=(4,))  # read below, never written

    def _alignment_heads(self):
        from ..word_timing import alignment_heads
        return self.options["alignment_heads"] or alignment_heads(self.model_name, self.dims["n_text_layer"], self.dims["n_text_head"])

    def transcribe_device(self, audio_dev: torch.Tensor, offs: np.ndarray, lens: np.ndarray, batch_size: int,
                          language: str, task: str, dtw_words: bool = False, chunk_starts=None):
        """mel -> encode -> greedy decode for chunks already resident in HBM.  Returns device tensors
        (tokens [n, sample_len], n_tokens, sum_logprob, no_speech_prob).  With dtw_words the decode kernel logs the
        alignment heads' cross-attention queries and every batch is followed by the DTW word timing (word_timing.py) while
        its cross-K cache is still resident: the result gains "words" (one list per chunk, times offset by chunk_starts)."""
        self._bind()
        self.ctx.collect_alignment_heads(self._alignment_heads() if dtw_words else None)
        words: List[list] = []
        without_ts = self.options["without_timestamps"]
        prompt = self.tokenizer.prompt(language, task, without_ts)
        n = len(offs)
        out = {"tokens": [], "n_tokens": [], "sum_logprob": [], "no_speech_prob": []}
        batch_size = int(batch_size)
        if not 1 <= batch_size <= MAX_BATCH:
            raise ValueError(f"batch_size={batch_size}: the decoder takes 1..{MAX_BATCH} sequences per call (include/wxb200.h)")
        for i in range(0, n, batch_size):
            j = min(n, i + batch_size)
            self.ctx.logmel_features(audio_dev, offs[i:j], lens[i:j], self.dims["n_mels"], self._filters)  # K1 -> K2 on the device
            enc = self.ctx.encode(None, n_chunks=j - i)
            r = self.ctx.decode_greedy(enc, prompt, self.specials["eot"], no_speech=self.specials["no_speech"],
                                       sample_len=int(self.options["sample_len"]),
                                       suppress_blank=bool(self.options["suppress_blank"]),
                                       blank_token=self.specials["blank"],
                                       suppress_tokens=tuple(self.options["suppress_tokens"]),
                                       timestamp_rules=None if without_ts else dict(
                                           timestamp_begin=self.specials["timestamp_begin"],
                                           no_timestamps=self.specials["no_timestamps"], max_initial_timestamp_index=50))
            for k in out:
                out[k].append(r[k])
            if dtw_words:
                from ..word_timing import dtw_word_timestamps
                toks, n_tok = r["tokens"].cpu().numpy(), r["n_tokens"].cpu().numpy()
                words += dtw_word_timestamps(self.ctx, [toks[k, : n_tok[k]] for k in range(j - i)], self.specials["eot"], len(prompt),
                                             lambda t: self.tokenizer.decode_piece(t),
                                             offsets=None if chunk_starts is None else chunk_starts[i:j])
        res = {k: torch.cat(v, 0) for k, v in out.items()}
        if dtw_words:
            res["words"] = words
        return res

    # ------------------------------------------------------------------ public API
    def transcribe_batch(self, segments: List[Dict[str, Any]], batch_size: int = 8, align_words: bool = False,
                         print_progress: bool = False, combined_progress: bool = False, verbose: bool = False,
                         language: Optional[str] = None, task: Optional[str] = None, dtw_words: Optional[bool] = None,
                         **kwargs) -> Dict[str, Any]:
        """`dtw_words` (default: asr_options["dtw_word_timestamps"]) adds "words" to every segment from the decoder's own
        cross-attention (the reference's single-stage word timing, mlx_whisper_optimized_final.py) — no alignment model."""
        dtw_words = bool(self.options["dtw_word_timestamps"] if dtw_words is None else dtw_words)
        segs = [s for s in segments if s.get("audio") is not None and len(s["audio"]) > 0]
        language = language or self.language
        if language is None:
            language = self.detect_language(segs[0]["audio"]) if segs else "en"
        task = task or self.task
        result_segments: List[Dict[str, Any]] = []
        if segs:
            audio_dev, offs, lens = self.upload_chunks([s["audio"] for s in segs])
            r = self.transcribe_device(audio_dev, offs, lens, max(1, int(batch_size or 8)), language, task, dtw_words=dtw_words,
                                       chunk_starts=[float(s["start"]) for s in segs])
            tokens = r["tokens"].cpu().numpy()          # D2H of the step's result
            n_tok = r["n_tokens"].cpu().numpy()
            sum_lp = r["sum_logprob"].cpu().numpy()
            nsp = r["no_speech_prob"].cpu().numpy()
            for k, seg in enumerate(segs):
                ids = tokens[k, : n_tok[k]].tolist()
                text = self.tokenizer.decode(ids).strip()
                if print_progress:
                    base = ((k + 1) / len(segs)) * 100
                    print(f"Progress: {(base / 2) if combined_progress else base:.2f}%...")
                if not text:
                    continue
                item = {"text": text, "start": round(float(seg["start"]), 3), "end": round(float(seg["end"]), 3),
                        "tokens": ids, "avg_logprob": float(sum_lp[k]) / (len(ids) + 1), "no_speech_prob": float(nsp[k])}
                if dtw_words:
                    item["words"] = r["words"][k]
                if verbose:
                    print(f"[{item['start']:.3f} --> {item['end']:.3f}] {text}")
                result_segments.append(item)
        result = {"segments": result_segments, "language": language or "en"}
        if align_words and result_segments:
            result = self._align_words(result, segments)
        return result

    def transcribe(self, audio: Union[str, np.ndarray], batch_size: Optional[int] = None, num_workers: int = 0,
                   language: Optional[str] = None, task: Optional[str] = None, chunk_size: int = 30,
                   print_progress: bool = False, combined_progress: bool = False, verbose: bool = False,
                   align_words: bool = False, **kwargs) -> TranscriptionResult:
        """Whole-audio entry point: cut into back-to-back <= chunk_size windows (the Lightning seek
        loop, mlx_lightning.py:174-216, without the per-window Python overhead) and batch them."""
        if isinstance(audio, str):
            audio = load_audio(audio)
        audio = np.asarray(audio, dtype=np.float32)
        win = int(min(chunk_size, 30) * SAMPLE_RATE)
        segments = []
        for s in range(0, max(len(audio), 1), win):
            e = min(len(audio), s + win)
            if e > s:
                segments.append({"start": s / SAMPLE_RATE, "end": e / SAMPLE_RATE, "audio": audio[s:e]})
        return self.transcribe_batch(segments, batch_size=batch_size or 8, align_words=align_words,
                                     print_progress=print_progress, combined_progress=combined_progress, verbose=verbose,
                                     language=language, task=task)

    def detect_language(self, audio: np.ndarray) -> str:
        """Language-id step: one decoder position on [sot]; argmax over the language tokens."""
        if not self.is_multilingual:
            return "en"
        self._bind()
        audio = np.asarray(audio, dtype=np.float32)[:N_SAMPLES]
        audio_dev, offs, lens = self.upload_chunks([audio])
        mel = self.ctx.logmel(audio_dev, offs, lens, N_SAMPLES, self.dims["n_mels"], self._filters)
        enc = self.ctx.encode(mel)
        logits = self.ctx.decoder_logits(enc, np.array([[self.specials["sot"]]], dtype=np.int32))[0, 0]
        lo = self.specials["sot"] + 1
        lang_tok = int(torch.argmax(logits[lo: lo + self.tokenizer.num_languages]).item()) + lo
        return self.tokenizer.language_of(lang_tok)

    # presence of this attribute makes the pipeline translate word_timestamps=True into
    # align_words=True (asr.py:50-52,76-78)
    def _align_words(self, result: Dict[str, Any], segments: List[Dict[str, Any]]) -> Dict[str, Any]:
        from ..alignment import align
        if self.align_model is None:
            raise RuntimeError("word alignment needs an alignment model: pass align_model=(model, metadata) to load_model "
                               "(whisperx.load_align_model needs network access for the default checkpoints)")
        model, meta = self.align_model
        segs = [s for s in segments if s.get("audio") is not None]
        if not segs:
            return result
        # stitch the chunk audio back onto one timeline for align()
        end = max(int(round(s["end"] * SAMPLE_RATE)) for s in segs)
        audio = np.zeros(end, dtype=np.float32)
        for s in segs:
            a = int(s["start"] * SAMPLE_RATE)
            audio[a:a + len(s["audio"])] = s["audio"][: max(0, end - a)]
        aligned = align(result["segments"], model, meta, audio, str(self.device))
        aligned["language"] = result["language"]
        return aligned
