"""K4 timing by mode and shape (CUDA events): trellis only / backtrack / beam-2, N tokens per segment, segments per launch.
usage (GPU box): python tools/ctc_probe.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisperx-mlx_b200")):
    sys.path.insert(0, p)
from whisperx._native import CTC_BACKTRACK, CTC_BEAM2, CTC_TRELLIS_ONLY, get_context  # noqa: E402

ctx = get_context(0)
T, V = 1499, 29


def timeit(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for n_seg in (15, 60):
    for N in (100, 450, 1040):
        rng = np.random.RandomState(N)
        em = torch.log_softmax(torch.randn(n_seg * T, V, device="cuda") * 3.0, -1)
        tok = torch.from_numpy(rng.randint(1, V, size=n_seg * N).astype(np.int32)).cuda()
        t_off = np.arange(n_seg + 1) * T
        n_off = np.arange(n_seg + 1) * N
        row = []
        for mode, nm in ((CTC_TRELLIS_ONLY, "trellis"), (CTC_BACKTRACK, "backtrack"), (CTC_BEAM2, "beam2")):
            row.append("%s %.3f ms" % (nm, timeit(lambda: ctx.ctc_align(em, t_off, tok, n_off, 0, mode))))
        print("segments %2d  N %4d  T %d:  %s" % (n_seg, N, T, "  ".join(row)), flush=True)
