"""
ctypes binding of libwxb200.so (include/wxb200.h) + a thin per-GPU context object.

PyTorch is used only to own device memory and streams; every numeric call below goes through the
C-ABI.  There is no CPU fallback: if the library is missing or no sm_100 GPU is present, the
constructors raise.
"""
import ctypes as C
import os
import threading
from typing import Dict, Optional

import numpy as np
import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "lib", "libwxb200.so")
_lib = None
_lib_lock = threading.Lock()

CTC_BACKTRACK, CTC_BEAM2, CTC_TRELLIS_ONLY = 0, 1, 2


def ctc_beam_mode(beam_width: int) -> int:
    """wxb_ctc_align mode of backtrack_beam with this beam width (WXB_CTC_BEAM(w) of include/wxb200.h)."""
    if not 1 <= int(beam_width) <= 8:
        raise NotImplementedError(f"backtrack_beam: beam_width={beam_width} (the kernel keeps 1..8 beams)")
    return CTC_BEAM2 | (int(beam_width) << 8)


class WxbError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_mels", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
        "n_vocab", "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer")]


class W2vDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("conv_dim", "embed_dim", "n_heads", "n_layers", "ff_dim", "pos_kernel", "pos_groups", "n_out")]


class VadParams(C.Structure):
    _fields_ = [("frame_duration", C.c_double), ("frame_step", C.c_double), ("frame_start", C.c_double), ("chunk_size", C.c_double),
                ("onset", C.c_float), ("offset", C.c_float)]


class DecodeOpts(C.Structure):
    _fields_ = [("eot", C.c_int32), ("no_speech", C.c_int32), ("sample_len", C.c_int32),
                ("suppress_blank", C.c_int32), ("blank_token", C.c_int32), ("n_suppress", C.c_int32),
                ("suppress_dev", C.c_void_p), ("check_every", C.c_int32), ("no_compaction", C.c_int32),
                ("apply_timestamp_rules", C.c_int32), ("timestamp_begin", C.c_int32), ("no_timestamps", C.c_int32),
                ("max_initial_timestamp_index", C.c_int32)]


def load_library(path: Optional[str] = None):
    """dlopen libwxb200.so and declare every prototype of include/wxb200.h."""
    global _lib
    with _lib_lock:
        if _lib is not None and path is None:
            return _lib
        p = path or os.environ.get("WXB200_LIB", _LIB_PATH)
        if path is None and "WXB200_LIB" in os.environ:
            import sys
            print(f"[whisperx b200] WXB200_LIB set: loading {p} instead of the in-tree library", file=sys.stderr)
        if not os.path.exists(p):
            raise WxbError(
                f"libwxb200.so not found at {p}: build it with `python __graft_entry__.py` "
                "(make -C whisperx-mlx_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(p)
        vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
        lib.wxb_abi_version.restype = i32
        lib.wxb_abi_version.argtypes = []
        lib.wxb_create.restype = i32
        lib.wxb_create.argtypes = [i32, C.POINTER(vp)]
        lib.wxb_destroy.restype = None
        lib.wxb_destroy.argtypes = [vp]
        lib.wxb_last_error.restype = C.c_char_p
        lib.wxb_last_error.argtypes = [vp]
        lib.wxb_launch_count.restype = i64
        lib.wxb_launch_count.argtypes = [vp]
        lib.wxb_logmel.restype = i32
        lib.wxb_logmel.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
        lib.wxb_logmel_features.restype = i32
        lib.wxb_logmel_features.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp]
        lib.wxb_ctc_align.restype = i32
        lib.wxb_ctc_align.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
        lib.wxb_log_softmax_rows.restype = i32
        lib.wxb_log_softmax_rows.argtypes = [vp, vp, i64, i32, vp]
        lib.wxb_set_model.restype = i32
        lib.wxb_set_model.argtypes = [vp, C.POINTER(Dims), C.POINTER(C.c_char_p), C.POINTER(vp), i32]
        lib.wxb_encode.restype = i32
        lib.wxb_encode.argtypes = [vp, vp, i32, vp, vp]
        lib.wxb_decode_greedy.restype = i32
        lib.wxb_decode_greedy.argtypes = [vp, vp, i32, vp, i32, C.POINTER(DecodeOpts), vp, vp, vp, vp, vp]
        lib.wxb_decode_stats.restype = i32
        lib.wxb_decode_stats.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), i32]
        lib.wxb_set_align_model.restype = i32
        lib.wxb_set_align_model.argtypes = [vp, C.POINTER(W2vDims), C.POINTER(C.c_char_p), C.POINTER(vp), i32]
        lib.wxb_w2v_frames.restype = i32
        lib.wxb_w2v_frames.argtypes = [i32]
        lib.wxb_w2v_emissions.restype = i32
        lib.wxb_w2v_emissions.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
        lib.wxb_debug_set.restype = i32
        lib.wxb_debug_set.argtypes = [vp, C.c_char_p, i32]
        lib.wxb_debug_copy.restype = i32
        lib.wxb_debug_copy.argtypes = [vp, C.c_char_p, vp, i64, i64]
        lib.wxb_decoder_sample.restype = i32
        lib.wxb_decoder_sample.argtypes = [vp, vp, i64, i32, i32, vp, i32, i32, i32, C.POINTER(DecodeOpts), vp, vp, vp, vp, vp]
        lib.wxb_decoder_logits.restype = i32
        lib.wxb_decoder_logits.argtypes = [vp, vp, i32, vp, i32, vp, vp]
        lib.wxb_gemm_bf16.restype = i32
        lib.wxb_gemm_bf16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
        lib.wxb_encoder_attention.restype = i32
        lib.wxb_encoder_attention.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
        lib.wxb_decode_collect_heads.restype = i32
        lib.wxb_decode_collect_heads.argtypes = [vp, vp, i32]
        lib.wxb_dtw_scores.restype = i32
        lib.wxb_dtw_scores.argtypes = [vp, i32, i32, vp, vp, vp]
        lib.wxb_dtw_cost.restype = i32
        lib.wxb_dtw_cost.argtypes = [vp, vp, i64, i32, C.c_float, i32, vp, vp]
        lib.wxb_dtw_path.restype = i32
        lib.wxb_dtw_path.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp]
        lib.wxb_dtw_path_capacity.restype = i32
        lib.wxb_dtw_path_capacity.argtypes = [i32]
        lib.wxb_vad_chunks.restype = i32
        lib.wxb_vad_chunks.argtypes = [vp, vp, vp, vp, i32, C.POINTER(VadParams), i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
        lib.wxb_vad_energy_scores.restype = i32
        lib.wxb_vad_energy_scores.argtypes = [vp, vp, i64, C.c_float, C.c_float, vp, vp]
        lib.wxb_vad_energy_frames.restype = i64
        lib.wxb_vad_energy_frames.argtypes = [i64]
        if lib.wxb_abi_version() != 3:
            raise WxbError("libwxb200.so ABI version mismatch")
        lib._wxb_path = os.path.abspath(p)
        if path is None:
            _lib = lib
        return lib


EXPORTED_SYMBOLS = (
    "wxb_abi_version", "wxb_create", "wxb_destroy", "wxb_last_error", "wxb_launch_count", "wxb_logmel",
    "wxb_ctc_align", "wxb_log_softmax_rows", "wxb_set_model", "wxb_encode", "wxb_decode_greedy",
    "wxb_decoder_logits", "wxb_gemm_bf16", "wxb_decode_stats", "wxb_encoder_attention", "wxb_decoder_sample",
    "wxb_logmel_features", "wxb_set_align_model", "wxb_w2v_frames", "wxb_w2v_emissions", "wxb_debug_set", "wxb_debug_copy",
    "wxb_vad_chunks", "wxb_vad_energy_scores", "wxb_vad_energy_frames", "wxb_decode_collect_heads", "wxb_dtw_scores", "wxb_dtw_cost", "wxb_dtw_path", "wxb_dtw_path_capacity")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One wxb_ctx per (process, GPU).  Not thread-safe (mirrors the C contract)."""

    def __init__(self, device_index: int = 0):
        if not torch.cuda.is_available():
            raise WxbError("whisperx b200 backend needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device_index = int(device_index)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        rc = self.lib.wxb_create(self.device_index, C.byref(h))
        if rc != 0:
            raise WxbError(self.lib.wxb_last_error(None).decode())
        self.h = h
        self._keep: Dict[str, object] = {}
        self.model_owner = None  # the object (backend) whose weight table is resident, see set_model
        self.align_owner = None  # same for the alignment model

    def close(self):
        if getattr(self, "h", None):
            self.lib.wxb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- helpers
    def _check(self, rc: int):
        if rc != 0:
            raise WxbError(f"[wxb {rc}] " + self.lib.wxb_last_error(self.h).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        return int(self.lib.wxb_launch_count(self.h))

    # ---------------------------------------------------------------- K1
    def logmel(self, audio_dev: torch.Tensor, chunk_off: np.ndarray, chunk_len: np.ndarray,
               n_samples_padded: int, n_mels: int, filters_dev: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio_dev f32 cuda [total]; returns f32 cuda [n_chunks, n_mels, n_samples_padded//160]."""
        assert audio_dev.is_cuda and audio_dev.dtype == torch.float32 and audio_dev.is_contiguous()
        assert filters_dev.is_cuda and filters_dev.dtype == torch.float32 and filters_dev.is_contiguous()
        off = np.ascontiguousarray(chunk_off, dtype=np.int64)
        ln = np.ascontiguousarray(chunk_len, dtype=np.int32)
        n = len(off)
        if n and int((off + ln).max()) > audio_dev.numel():
            raise ValueError("chunk exceeds the audio buffer")
        n_frames = n_samples_padded // 160
        if out is None:
            out = torch.empty((n, n_mels, n_frames), dtype=torch.float32, device=self.device)
        assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == n * n_mels * n_frames
        self._check(self.lib.wxb_logmel(self.h, _ptr(audio_dev), _np_ptr(off), _np_ptr(ln), n, n_samples_padded,
                                        n_mels, _ptr(filters_dev), _ptr(out), self._stream()))
        return out

    def logmel_features(self, audio_dev: torch.Tensor, chunk_off: np.ndarray, chunk_len: np.ndarray, n_mels: int,
                        filters_dev: torch.Tensor, want_f32: bool = False) -> Optional[torch.Tensor]:
        """K1 -> K2 hand-off for 30 s chunks: the log-mel is left in the context's encoder input buffer (bf16, frame-major);
        follow with encode(None, n_chunks=...).  Returns the f32 [n, n_mels, 3000] tensor only if want_f32."""
        assert audio_dev.is_cuda and audio_dev.dtype == torch.float32 and audio_dev.is_contiguous()
        off = np.ascontiguousarray(chunk_off, dtype=np.int64)
        ln = np.ascontiguousarray(chunk_len, dtype=np.int32)
        n = len(off)
        if n and int((off + ln).max()) > audio_dev.numel():
            raise ValueError("chunk exceeds the audio buffer")
        out = torch.empty((n, n_mels, 3000), dtype=torch.float32, device=self.device) if want_f32 else None
        self._check(self.lib.wxb_logmel_features(self.h, _ptr(audio_dev), _np_ptr(off), _np_ptr(ln), n, n_mels, _ptr(filters_dev),
                                                 _ptr(out), self._stream()))
        return out

    # ---------------------------------------------------------------- K4
    def ctc_align(self, emis_dev: torch.Tensor, t_off: np.ndarray, tok_dev: torch.Tensor, n_off: np.ndarray,
                  blank: int, mode: int, want_trellis: bool = False):
        """emis_dev f32 cuda [sumT, V] log-probs.  Returns dict of cuda tensors."""
        assert emis_dev.is_cuda and emis_dev.dtype == torch.float32 and emis_dev.is_contiguous() and emis_dev.dim() == 2
        assert tok_dev.is_cuda and tok_dev.dtype == torch.int32 and tok_dev.is_contiguous()
        t_off = np.ascontiguousarray(t_off, dtype=np.int32)
        n_off = np.ascontiguousarray(n_off, dtype=np.int32)
        n_seg = len(t_off) - 1
        sumT, V = emis_dev.shape
        assert int(t_off[-1]) == sumT and int(n_off[-1]) == tok_dev.numel()
        dev = self.device
        path_tok = torch.full((max(sumT, 1),), -1, dtype=torch.int32, device=dev)
        path_lp = torch.zeros((max(sumT, 1),), dtype=torch.float32, device=dev)
        path_prob = torch.zeros((max(sumT, 1),), dtype=torch.float32, device=dev)
        status = torch.full((max(n_seg, 1),), -1, dtype=torch.int32, device=dev)
        trellis = None
        if want_trellis:
            T = np.diff(t_off).astype(np.int64)
            N = np.diff(n_off).astype(np.int64)
            trellis = torch.empty((max(int((T * N).sum()), 1),), dtype=torch.float32, device=dev)
        self._check(self.lib.wxb_ctc_align(self.h, _ptr(emis_dev), _np_ptr(t_off), _ptr(tok_dev), _np_ptr(n_off),
                                           n_seg, V, int(blank), int(mode), _ptr(trellis), _ptr(path_tok),
                                           _ptr(path_lp), _ptr(path_prob), _ptr(status), self._stream()))
        return dict(path_tok=path_tok[:sumT], path_lp=path_lp[:sumT], path_prob=path_prob[:sumT],
                    status=status[:n_seg], trellis=trellis)

    def log_softmax_rows_(self, x: torch.Tensor) -> torch.Tensor:
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        V = x.shape[-1]
        self._check(self.lib.wxb_log_softmax_rows(self.h, _ptr(x), x.numel() // V, V, self._stream()))
        return x

    # ---------------------------------------------------------------- K2 / K3
    def set_model(self, dims: dict, tensors: Dict[str, torch.Tensor], owner=None):
        """Borrow `tensors` as the resident model.  `owner` tags whose model it is: a Context is shared by everything on
        its GPU, so callers that keep a model across calls re-bind when the owner changed (B200WhisperBackend._bind)."""
        d = Dims(**{k: int(dims[k]) for k, _ in Dims._fields_})
        names = list(tensors.keys())
        for k in names:
            t = tensors[k]
            assert t.is_cuda and t.is_contiguous(), k
        arr_n = (C.c_char_p * len(names))(*[n.encode() for n in names])
        arr_p = (C.c_void_p * len(names))(*[tensors[n].data_ptr() for n in names])
        self._keep["model"] = tensors  # the library borrows the pointers
        self._keep["dims"] = dict(dims)
        self.model_owner = None
        self._check(self.lib.wxb_set_model(self.h, C.byref(d), arr_n, arr_p, len(names)))
        self.model_owner = owner

    def set_align_model(self, dims: dict, tensors: Dict[str, torch.Tensor], owner=None):
        d = W2vDims(**{k: int(dims[k]) for k, _ in W2vDims._fields_})
        names = list(tensors.keys())
        for k in names:
            assert tensors[k].is_cuda and tensors[k].is_contiguous(), k
        arr_n = (C.c_char_p * len(names))(*[n.encode() for n in names])
        arr_p = (C.c_void_p * len(names))(*[tensors[n].data_ptr() for n in names])
        self._keep["align_model"] = tensors  # the library borrows the pointers
        self.align_owner = None
        self._check(self.lib.wxb_set_align_model(self.h, C.byref(d), arr_n, arr_p, len(names)))
        self.align_owner = owner

    def w2v_emissions(self, audio_dev: torch.Tensor, seg_off: np.ndarray, seg_len: np.ndarray, emis_out: torch.Tensor,
                      t_off: np.ndarray):
        assert audio_dev.is_cuda and audio_dev.dtype == torch.float32 and audio_dev.is_contiguous()
        assert emis_out.is_cuda and emis_out.dtype == torch.float32 and emis_out.is_contiguous()
        off = np.ascontiguousarray(seg_off, dtype=np.int64)
        ln = np.ascontiguousarray(seg_len, dtype=np.int32)
        to = np.ascontiguousarray(t_off, dtype=np.int32)
        if len(off) and int((off + ln).max()) > audio_dev.numel():
            raise ValueError("segment exceeds the audio buffer")
        self._check(self.lib.wxb_w2v_emissions(self.h, _ptr(audio_dev), _np_ptr(off), _np_ptr(ln), len(off), _ptr(emis_out),
                                               _np_ptr(to), self._stream()))
        return emis_out

    def debug_set(self, key: str, value: int):
        self._check(self.lib.wxb_debug_set(self.h, key.encode(), int(value)))

    def debug_buffer(self, name: str, shape, dtype, offset_bytes: int = 0) -> torch.Tensor:
        out = torch.empty(shape, dtype=dtype, device=self.device)
        self._check(self.lib.wxb_debug_copy(self.h, name.encode(), _ptr(out), int(offset_bytes), out.numel() * out.element_size()))
        return out

    def encode(self, mel_dev: Optional[torch.Tensor], n_chunks: Optional[int] = None) -> torch.Tensor:
        """mel f32 cuda [B, n_mels, 3000] -> bf16 cuda [B, 1500, d].  mel_dev = None: encode the n_chunks chunks whose
        log-mel the last logmel_features call left on the device."""
        dims = self._keep["dims"]
        if mel_dev is None:
            B = int(n_chunks)
        else:
            assert mel_dev.is_cuda and mel_dev.dtype == torch.float32 and mel_dev.is_contiguous()
            B = mel_dev.shape[0]
        out = torch.empty((B, dims["n_audio_ctx"], dims["n_audio_state"]), dtype=torch.bfloat16, device=self.device)
        self._check(self.lib.wxb_encode(self.h, _ptr(mel_dev), B, _ptr(out), self._stream()))
        return out

    def decode_greedy(self, enc_out: torch.Tensor, prompt, eot: int, no_speech: int = -1, sample_len: int = 224,
                      suppress_blank: bool = False, blank_token: int = 220, suppress_tokens=(),
                      check_every: int = 16, compaction: bool = True, timestamp_rules: Optional[dict] = None):
        """timestamp_rules = dict(timestamp_begin=, no_timestamps=, max_initial_timestamp_index=) switches
        ApplyTimestampRules on (decode with timestamps); None = off (prompt ends with <|notimestamps|>)."""
        assert enc_out.is_cuda and enc_out.dtype == torch.bfloat16 and enc_out.is_contiguous()
        B = enc_out.shape[0]
        dev = self.device
        prompt_np = np.ascontiguousarray(prompt, dtype=np.int32)
        opts, sup = self._decode_opts(eot, no_speech, sample_len, suppress_blank, blank_token, suppress_tokens, check_every,
                                      compaction, timestamp_rules)
        tokens = torch.full((B, sample_len), eot, dtype=torch.int32, device=dev)
        n_tok = torch.zeros((B,), dtype=torch.int32, device=dev)
        sum_lp = torch.zeros((B,), dtype=torch.float32, device=dev)
        nsp = torch.zeros((B,), dtype=torch.float32, device=dev)
        self._check(self.lib.wxb_decode_greedy(self.h, _ptr(enc_out), B, _np_ptr(prompt_np), len(prompt_np),
                                               C.byref(opts), _ptr(tokens), _ptr(n_tok), _ptr(sum_lp), _ptr(nsp),
                                               self._stream()))
        return dict(tokens=tokens, n_tokens=n_tok, sum_logprob=sum_lp, no_speech_prob=nsp)

    def _decode_opts(self, eot, no_speech, sample_len, suppress_blank, blank_token, suppress_tokens, check_every, compaction,
                     timestamp_rules):
        dev = self.device
        sup = torch.tensor(list(suppress_tokens), dtype=torch.int32, device=dev) if len(suppress_tokens) else None
        opts = DecodeOpts(eot=eot, no_speech=no_speech, sample_len=sample_len, suppress_blank=int(suppress_blank),
                          blank_token=blank_token, n_suppress=0 if sup is None else sup.numel(),
                          suppress_dev=None if sup is None else sup.data_ptr(), check_every=check_every,
                          no_compaction=0 if compaction else 1)
        if timestamp_rules:
            opts.apply_timestamp_rules = 1
            opts.timestamp_begin = int(timestamp_rules["timestamp_begin"])
            opts.no_timestamps = int(timestamp_rules.get("no_timestamps", -1))
            mi = timestamp_rules.get("max_initial_timestamp_index", 50)
            opts.max_initial_timestamp_index = -1 if mi is None else int(mi)
        return opts, sup

    def sample_step(self, logits: torch.Tensor, tokens: torch.Tensor, pos: int, prompt_len: int, eot: int,
                    sum_logprob: torch.Tensor, done: torch.Tensor, ts_last: Optional[torch.Tensor] = None,
                    no_speech: int = -1, suppress_blank: bool = False, blank_token: int = 220, suppress_tokens=(),
                    timestamp_rules: Optional[dict] = None):
        """The decoder's sampling phase on caller-supplied logits (parity tests of the greedy update rule).
        logits f32 cuda [B, V] (copied into a 16-byte-aligned padded buffer); tokens int32 cuda [B, stride] updated in
        place at column pos + 1; sum_logprob f32 [B], done int32 [B], ts_last int32 [B] updated in place."""
        B, V = logits.shape
        ldl = (V + 3) & ~3
        buf = torch.full((B, ldl), float("-inf"), dtype=torch.float32, device=self.device)
        buf[:, :V] = logits
        opts, sup = self._decode_opts(eot, no_speech, 1, suppress_blank, blank_token, suppress_tokens, 16, True, timestamp_rules)
        nsp = torch.zeros((B,), dtype=torch.float32, device=self.device) if no_speech >= 0 else None
        assert tokens.dtype == torch.int32 and tokens.is_contiguous() and done.dtype == torch.int32
        self._check(self.lib.wxb_decoder_sample(self.h, _ptr(buf), ldl, B, V, _ptr(tokens), tokens.shape[1], int(pos),
                                                int(prompt_len), C.byref(opts), _ptr(sum_logprob), _ptr(done), _ptr(ts_last),
                                                _ptr(nsp), self._stream()))
        return buf[:, :V], nsp

    def decode_stats(self, reset: bool = True):
        """(cross_kv_ms, steps_ms, n_steps) summed over the decode_greedy calls since the last reset."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.wxb_decode_stats(self.h, C.byref(a), C.byref(b), C.byref(n), int(reset)))
        return a.value, b.value, n.value

    def decoder_logits(self, enc_out: torch.Tensor, tokens: np.ndarray) -> torch.Tensor:
        dims = self._keep["dims"]
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        B, n = tokens.shape
        out = torch.empty((B, n, dims["n_vocab"]), dtype=torch.float32, device=self.device)
        self._check(self.lib.wxb_decoder_logits(self.h, _ptr(enc_out), B, _np_ptr(tokens), n, _ptr(out), self._stream()))
        return out

    # ---- word timing from cross-attention (wxb_dtw.cu) -----------------------------------------------------------------
    def collect_alignment_heads(self, heads):
        """(layer, head) pairs whose cross-attention queries later decodes log on the device; None / [] = off."""
        h = np.ascontiguousarray(np.asarray(heads if heads is not None else [], dtype=np.int32).reshape(-1, 2))
        self._check(self.lib.wxb_decode_collect_heads(self.h, _np_ptr(h) if len(h) else None, len(h)))

    def dtw_scores(self, n_rows: np.ndarray, pos0: int, n_frames: int = 1500) -> torch.Tensor:
        """Mean-over-alignment-heads pre-softmax cross-attention scores of the last decode: f32 [sum n_rows, n_frames]."""
        n_rows = np.ascontiguousarray(n_rows, dtype=np.int32)
        out = torch.empty((int(n_rows.sum()), n_frames), dtype=torch.float32, device=self.device)
        self._check(self.lib.wxb_dtw_scores(self.h, len(n_rows), int(pos0), _np_ptr(n_rows), _ptr(out), self._stream()))
        return out

    def dtw_cost(self, qk: torch.Tensor, temperature: float = 10.0, medfilt_width: int = 7) -> torch.Tensor:
        assert qk.dtype == torch.float32 and qk.is_contiguous() and qk.dim() == 2
        out = torch.empty_like(qk)
        self._check(self.lib.wxb_dtw_cost(self.h, _ptr(qk), qk.shape[0], qk.shape[1], float(temperature), int(medfilt_width),
                                          _ptr(out), self._stream()))
        return out

    def dtw_path(self, cost: torch.Tensor, n_rows: np.ndarray):
        """cost f32 [sum n_rows, T] (token rows of every sequence back to back) -> list of int32 arrays [2, P_b]
        (row 0 frame indices, row 1 token indices: the reference's dtw(-w.T) result)."""
        assert cost.dtype == torch.float32 and cost.is_contiguous() and cost.dim() == 2
        n_rows = np.ascontiguousarray(n_rows, dtype=np.int32)
        assert int(n_rows.sum()) == cost.shape[0]
        B, T = len(n_rows), cost.shape[1]
        cap = self.lib.wxb_dtw_path_capacity(T)
        pf = torch.empty((B, cap), dtype=torch.int32, device=self.device)
        pt = torch.empty((B, cap), dtype=torch.int32, device=self.device)
        pl = torch.empty((B,), dtype=torch.int32, device=self.device)
        self._check(self.lib.wxb_dtw_path(self.h, _ptr(cost), B, _np_ptr(n_rows), T, _ptr(pf), _ptr(pt), _ptr(pl), self._stream()))
        pl_h = pl.cpu().numpy()
        n_max = int(pl_h.max()) if B else 0
        pf_h, pt_h = pf[:, :n_max].cpu().numpy(), pt[:, :n_max].cpu().numpy()
        return [np.stack([pf_h[b, :pl_h[b]], pt_h[b, :pl_h[b]]]) for b in range(B)]

    # ---- VAD post-processing and chunking (wxb_vad.cu) ------------------------------------------------------------------
    def vad_energy_scores(self, audio_dev: torch.Tensor, floor_db: float = -50.0, width_db: float = 6.0) -> torch.Tensor:
        assert audio_dev.is_cuda and audio_dev.dtype == torch.float32 and audio_dev.is_contiguous() and audio_dev.dim() == 1
        out = torch.empty((int(self.lib.wxb_vad_energy_frames(audio_dev.numel())),), dtype=torch.float32, device=self.device)
        self._check(self.lib.wxb_vad_energy_scores(self.h, _ptr(audio_dev), audio_dev.numel(), float(floor_db), float(width_db),
                                                   _ptr(out), self._stream()))
        return out

    def vad_chunks(self, scores_dev: torch.Tensor, score_off: np.ndarray, n_samples: np.ndarray, chunk_size: float, onset: float = 0.5,
                   offset: Optional[float] = None, frame_duration: float = 0.0619375, frame_step: float = 0.016875,
                   frame_start: float = 0.0, max_regions: int = 0, max_chunks: int = 0):
        """Binarize + merge_chunks for the score tracks of len(score_off) - 1 recordings.  Returns per recording
        dict(regions f64 [n, 2], chunks f64 [m, 2], chunk_first int32 [m + 1], chunk_off int64 [m], chunk_len int32 [m]) with the
        chunk_off / chunk_len tables also left on the device under "chunk_off_dev" / "chunk_len_dev"."""
        assert scores_dev.is_cuda and scores_dev.dtype == torch.float32 and scores_dev.is_contiguous()
        off = np.ascontiguousarray(score_off, dtype=np.int64)
        ns = np.ascontiguousarray(n_samples, dtype=np.int64)
        n_rec = len(off) - 1
        longest = int(np.diff(off).max()) if n_rec else 0
        prm = VadParams(float(frame_duration), float(frame_step), float(frame_start), float(chunk_size), float(onset),
                        float(offset) if offset else 0.0)
        max_regions = int(max_regions) or max(16, longest // 2 + 2)  # a region needs at least two frames
        max_chunks = int(max_chunks) or max_regions
        dev = self.device
        regions = torch.empty((n_rec, max_regions, 2), dtype=torch.float64, device=dev)
        chunks = torch.empty((n_rec, max_chunks, 2), dtype=torch.float64, device=dev)
        first = torch.empty((n_rec, max_chunks + 1), dtype=torch.int32, device=dev)
        n_reg = torch.empty((n_rec,), dtype=torch.int32, device=dev)
        n_ch = torch.empty((n_rec,), dtype=torch.int32, device=dev)
        c_off = torch.empty((n_rec, max_chunks), dtype=torch.int64, device=dev)
        c_len = torch.empty((n_rec, max_chunks), dtype=torch.int32, device=dev)
        self._check(self.lib.wxb_vad_chunks(self.h, _ptr(scores_dev), _np_ptr(off), _np_ptr(ns), n_rec, C.byref(prm), max_regions,
                                            max_chunks, _ptr(regions), _ptr(n_reg), _ptr(chunks), _ptr(first), _ptr(n_ch),
                                            _ptr(c_off), _ptr(c_len), self._stream()))
        n_reg_h, n_ch_h = n_reg.cpu().numpy(), n_ch.cpu().numpy()
        if (n_reg_h > max_regions).any() or (n_ch_h > max_chunks).any():
            raise WxbError(f"vad_chunks: {int(n_reg_h.max())} regions / {int(n_ch_h.max())} chunks exceed the capacities "
                           f"({max_regions}, {max_chunks})")
        mr, mc = int(n_reg_h.max()) if n_rec else 0, int(n_ch_h.max()) if n_rec else 0
        regions_h, chunks_h = regions[:, :mr].cpu().numpy(), chunks[:, :mc].cpu().numpy()
        first_h, off_h, len_h = first[:, :mc + 1].cpu().numpy(), c_off[:, :mc].cpu().numpy(), c_len[:, :mc].cpu().numpy()
        out = []
        for r in range(n_rec):
            m = int(n_ch_h[r])
            fr = first_h[r, :m + 1].copy()
            fr[m] = n_reg_h[r]
            out.append(dict(regions=regions_h[r, :n_reg_h[r]], chunks=chunks_h[r, :m], chunk_first=fr, chunk_off=off_h[r, :m],
                            chunk_len=len_h[r, :m], chunk_off_dev=c_off[r, :m], chunk_len_dev=c_len[r, :m]))
        return out

    def gemm_bf16(self, A: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, gelu: bool = False,
                  out_f32: bool = False) -> torch.Tensor:
        assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16 and A.is_contiguous() and W.is_contiguous()
        M, K = A.shape
        N, K2 = W.shape
        assert K == K2
        D = torch.empty((M, N), dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
        flags = (1 if gelu else 0) | (2 if out_f32 else 0)
        self._check(self.lib.wxb_gemm_bf16(self.h, _ptr(A), _ptr(W), _ptr(bias), _ptr(D), M, N, K, flags, self._stream()))
        return D


    def encoder_attention(self, qkv: torch.Tensor, B: int, T: int, H: int) -> torch.Tensor:
        """qkv bf16 [B*T, 3*64*H] -> softmax(Q K^T / 8) V per head, bf16 [B*T, 64*H]."""
        d = 64 * H
        assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and tuple(qkv.shape) == (B * T, 3 * d)
        out = torch.empty((B * T, d), dtype=torch.bfloat16, device=self.device)
        self._check(self.lib.wxb_encoder_attention(self.h, _ptr(qkv), _ptr(out), B, T, d, H, self._stream()))
        return out


_contexts: Dict[int, Context] = {}


def get_context(device_index: int = 0) -> Context:
    """Process-wide context per GPU (the backend, audio.log_mel_spectrogram and align() share it)."""
    ctx = _contexts.get(device_index)
    if ctx is None:
        ctx = Context(device_index)
        _contexts[device_index] = ctx
    return ctx
