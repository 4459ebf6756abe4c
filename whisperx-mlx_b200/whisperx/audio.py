"""
Audio frontend of the B200 backend.  Same names/constants as the reference's whisperx/audio.py;
`log_mel_spectrogram` runs kernel K1 (csrc/wxb_logmel.cu) through the C-ABI.
"""
import os
import subprocess
from functools import lru_cache
from typing import Optional, Sequence, Union

import numpy as np
import torch

from .utils import exact_div

# whisperx/audio.py:13-22
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_FRAMES = exact_div(N_SAMPLES, HOP_LENGTH)
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = exact_div(SAMPLE_RATE, HOP_LENGTH)
TOKENS_PER_SECOND = exact_div(SAMPLE_RATE, N_SAMPLES_PER_TOKEN)


def load_audio(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """Decode `file` to mono f32 at `sr` Hz by piping through the ffmpeg CLI (whisperx/audio.py:25-65)."""
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1",
           "-acodec", "pcm_s16le", "-ar", str(sr), "-"]
    try:
        pcm = subprocess.run(cmd, capture_output=True, check=True).stdout
    except subprocess.CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    except FileNotFoundError as e:
        raise RuntimeError("Failed to load audio: the ffmpeg CLI is not installed") from e
    return np.frombuffer(pcm, np.int16).flatten().astype(np.float32) / 32768.0


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Cut or zero-pad `array` along `axis` to exactly `length` (whisperx/audio.py:68-91)."""
    n = array.shape[axis]
    if torch.is_tensor(array):
        if n > length:
            array = array.narrow(axis, 0, length)
        elif n < length:
            shape = list(array.shape)
            shape[axis] = length - n
            array = torch.cat([array, array.new_zeros(shape)], dim=axis)
        return array
    if n > length:
        array = np.take(array, np.arange(length), axis=axis)
    elif n < length:
        widths = [(0, 0)] * array.ndim
        widths[axis] = (0, length - n)
        array = np.pad(array, widths)
    return array


@lru_cache(maxsize=None)
def _mel_filters_np(n_mels: int) -> np.ndarray:
    path = os.path.join(os.path.dirname(__file__), "assets", "mel_filters.npz")
    with np.load(path) as f:
        return np.ascontiguousarray(f[f"mel_{n_mels}"], dtype=np.float32)


@lru_cache(maxsize=None)
def mel_filters(device, n_mels: int) -> torch.Tensor:
    """Mel filterbank [n_mels, 201] (whisperx/audio.py:94-109); data file assets/mel_filters.npz."""
    assert n_mels in (80, 128), f"Unsupported n_mels: {n_mels}"
    return torch.from_numpy(_mel_filters_np(n_mels)).to(device)


def _cuda_index(device) -> int:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise RuntimeError("the b200 backend computes log-mel on the GPU only (device must be cuda); no CPU fallback")
    return dev.index if dev.index is not None else torch.cuda.current_device()


def log_mel_spectrogram(audio: Union[str, np.ndarray, torch.Tensor], n_mels: int, padding: int = 0,
                        device: Optional[Union[str, torch.device]] = None) -> torch.Tensor:
    """Drop-in for whisperx/audio.py:112-159: f32 [n_mels, (S+padding)//160] on the GPU."""
    if isinstance(audio, str):
        audio = load_audio(audio)
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32))
    if audio.dim() != 1:
        raise ValueError("log_mel_spectrogram expects a 1-D waveform (the hot path calls it per chunk)")
    if device is None and audio.is_cuda:
        device = audio.device
    idx = _cuda_index(device)
    from ._native import get_context
    ctx = get_context(idx)
    x = audio.to(ctx.device, dtype=torch.float32).contiguous()
    total = x.numel() + int(padding)
    n_frames = total // HOP_LENGTH
    out = ctx.logmel(x, np.array([0]), np.array([x.numel()]), total, n_mels, mel_filters(ctx.device, n_mels))
    return out[0, :, :n_frames]


def log_mel_chunks(chunks: Sequence[np.ndarray], n_mels: int, device_index: int = 0,
                   n_samples: int = N_SAMPLES) -> torch.Tensor:
    """Hot-path batch form: every chunk zero-padded to `n_samples`, one launch for all chunks.
    Returns f32 cuda [n_chunks, n_mels, n_samples//160]."""
    from ._native import get_context
    ctx = get_context(device_index)
    lens = np.array([min(len(c), n_samples) for c in chunks], dtype=np.int32)
    offs = np.zeros(len(chunks), dtype=np.int64)
    if len(chunks) > 1:
        offs[1:] = np.cumsum(lens[:-1])
    host = torch.empty(int(lens.sum()), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for c, o, l in zip(chunks, offs, lens):
        hv[o:o + l] = np.asarray(c[:l], dtype=np.float32)
    dev = host.to(ctx.device, non_blocking=True)
    return ctx.logmel(dev, offs, lens, n_samples, n_mels, mel_filters(ctx.device, n_mels))
