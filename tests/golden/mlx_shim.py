"""
A ~60-line numpy stand-in for `mlx.core` (+ empty stand-ins for the un-vendored `mlx_whisper` modules), just enough to
IMPORT AND RUN the reference's own in-tree decoding code on the CPU of the build container:

    /root/reference/mlx_whisper_batch_decoder.py:262-303   BatchGreedyDecoder.update
    /root/reference/mlx_whisper_batch_decoder.py:317-384   BatchDecodingTask._main_loop_batch
    /root/reference/mlx_ultra_optimized_batch.py:38-71     the batch-safe ApplyTimestampRules clause

Used only by tests/golden/make_greedy_golden.py (test infrastructure; never imported by the product).
"""
import sys
import types

import numpy as np


class array(np.ndarray):
    """ndarray with the two mlx-only methods the reference calls."""

    def logsumexp(self, axis=None, keepdims=False):
        return logsumexp(self, axis=axis, keepdims=keepdims)

    def astype(self, dtype, *a, **k):  # mx.array.astype(mx.float32)
        return np.ndarray.astype(self, dtype, *a, **k).view(array)


def _wrap(x):
    return np.asarray(x).view(array)


def logsumexp(x, axis=None, keepdims=False):
    x = np.asarray(x, dtype=np.float32)
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0).astype(np.float32)
    out = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True, dtype=np.float32)) + m
    return _wrap(out if keepdims else np.squeeze(out, axis=axis))


def softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float32)
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return _wrap(e / np.sum(e, axis=axis, keepdims=True, dtype=np.float32))


def install():
    mx = types.ModuleType("mlx.core")
    mx.array = lambda x, dtype=None: _wrap(np.array(x, dtype=dtype))
    mx.zeros = lambda shape, dtype=np.float32: _wrap(np.zeros(shape, dtype=dtype))
    mx.ones = lambda shape, dtype=np.float32: _wrap(np.ones(shape, dtype=dtype))
    mx.ones_like = lambda x: _wrap(np.ones_like(x))
    mx.full = lambda shape, v, dtype=np.float32: _wrap(np.full(shape, v, dtype=dtype))
    mx.arange = lambda *a, **k: _wrap(np.arange(*a, **k))
    mx.where = lambda c, a, b: _wrap(np.where(c, a, b))
    mx.concatenate = lambda xs, axis=0: _wrap(np.concatenate(xs, axis=axis))
    mx.take = lambda x, idx, axis=0: _wrap(np.take(x, idx, axis=axis))
    mx.broadcast_to = lambda x, shape: _wrap(np.broadcast_to(x, shape))
    mx.all = lambda x: bool(np.all(x))
    mx.logsumexp, mx.softmax = logsumexp, softmax
    mx.eval = lambda *a: None
    mx.float32, mx.float16, mx.int32, mx.bool_, mx.nan = np.float32, np.float16, np.int32, np.bool_, np.nan
    mlx = types.ModuleType("mlx")
    mlx.core = mx
    nn = types.ModuleType("mlx.nn")
    mlx.nn = nn
    sys.modules.update({"mlx": mlx, "mlx.core": mx, "mlx.nn": nn})

    # mlx_whisper is un-vendored (pyproject.toml:13): only the NAMES the reference files import are provided
    mw = types.ModuleType("mlx_whisper")
    dec = types.ModuleType("mlx_whisper.decoding")

    class GreedyDecoder:
        def __init__(self, temperature, eot):
            self.temperature, self.eot = temperature, eot

    class DecodingTask:
        pass

    class ApplyTimestampRules:
        def __init__(self, tokenizer):
            self.tokenizer = tokenizer

        def apply(self, logits, tokens):  # replaced by the reference's install_broadcasting_fix()
            raise NotImplementedError

    for name in ("DecodingOptions", "DecodingResult", "TokenDecoder", "LogitFilter"):
        setattr(dec, name, type(name, (), {}))
    dec.GreedyDecoder, dec.DecodingTask, dec.ApplyTimestampRules = GreedyDecoder, DecodingTask, ApplyTimestampRules
    mods = {"mlx_whisper": mw, "mlx_whisper.decoding": dec}
    for sub, names in (("tokenizer", ("Tokenizer", "get_tokenizer")), ("load_models", ("load_model",)),
                       ("audio", ("log_mel_spectrogram", "pad_or_trim")), ("timing", ("add_word_timestamps",))):
        m = types.ModuleType("mlx_whisper." + sub)
        for n in names:
            setattr(m, n, type(n, (), {}))
        setattr(mw, sub, m)
        mods["mlx_whisper." + sub] = m
    mw.decoding = dec
    sys.modules.update(mods)
    return mx
