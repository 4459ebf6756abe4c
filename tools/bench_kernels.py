"""Per-kernel timing with CUDA events (warm-up, L2 flush between iterations)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisperx-mlx_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from whisperx._native import CTC_BACKTRACK, CTC_BEAM2, CTC_TRELLIS_ONLY, get_context  # noqa: E402
import whisperx.audio as wa  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ctx = get_context(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    # ---- K1: 60 x 30 s chunks, 128 mels
    for n_chunks, n_mels in ((60, 128), (8, 128), (16, 80)):
        audio = torch.randn(n_chunks * 480000, device="cuda") * 0.1
        offs = np.arange(n_chunks, dtype=np.int64) * 480000
        lens = np.full(n_chunks, 480000, np.int32)
        filt = wa.mel_filters(ctx.device, n_mels)
        mel = torch.empty((n_chunks, n_mels, 3000), device="cuda")
        med, best = timeit(lambda: ctx.logmel(audio, offs, lens, 480000, n_mels, filt, out=mel), flush=flush)
        byts = n_chunks * (4 * 480000 + n_mels * 3000 * 4)
        out[f"logmel_{n_chunks}x{n_mels}"] = dict(ms=med, ms_best=best, gbs=byts / med / 1e6, frac=byts / med / 1e6 / PEAKS["hbm_gbs"])
    # ---- K4: 60 segments T=1499, N~U(50,450)
    rng = np.random.RandomState(0)
    n_seg, T, V = 60, 1499, 29
    Ns = rng.randint(50, 451, size=n_seg)
    em = torch.log_softmax(torch.randn(n_seg * T, V, device="cuda"), -1)
    tok = torch.from_numpy(rng.randint(1, V, size=int(Ns.sum())).astype(np.int32)).cuda()
    t_off = np.arange(n_seg + 1) * T
    n_off = np.concatenate([[0], np.cumsum(Ns)])
    byts = float(sum(4 * T * V + 4 * T * n + 16 * T for n in Ns))
    for mode, nm in ((CTC_TRELLIS_ONLY, "trellis"), (CTC_BACKTRACK, "backtrack"), (CTC_BEAM2, "beam2")):
        med, best = timeit(lambda: ctx.ctc_align(em, t_off, tok, n_off, 0, mode), flush=flush)
        out[f"ctc_{nm}_{n_seg}seg"] = dict(ms=med, ms_best=best, gbs=byts / med / 1e6)
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_kernels.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
