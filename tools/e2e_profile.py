"""Where the end-to-end (host API) time goes: cProfile of bench.py's e2e step (load_model -> transcribe_sharded -> whisperx.align).
usage (GPU box): python tools/e2e_profile.py [n_steps]"""
import cProfile
import os
import pstats
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "whisperx-mlx_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import bench  # noqa: E402
import whisperx  # noqa: E402
from whisperx import multi_gpu  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
audio = bench.job_audio(30.0, 1234)
bundle = whisperx.load_align_model("en", dev, model_name="WAV2VEC2_ASR_BASE_960H", random_init=True)
pipe = whisperx.load_model("large-v3", device="cuda", backend="b200", language="en", vad_method="uniform", batch_size=60, align_model=bundle)
sections = {}


def align_fn(local, mine):
    t0 = time.perf_counter()
    out = whisperx.align(local["segments"], bundle[0], bundle[1], audio, str(dev))
    torch.cuda.synchronize()
    sections.setdefault("align", []).append(time.perf_counter() - t0)
    out["language"] = local["language"]
    return out


def step():
    t0 = time.perf_counter()
    out = multi_gpu.transcribe_sharded(pipe, audio, 0, 1, batch_size=60, chunk_size=30, align_fn=align_fn)
    torch.cuda.synchronize()
    sections.setdefault("total", []).append(time.perf_counter() - t0)
    return out


for _ in range(2):
    step()
sections.clear()
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
pr.disable()
print({k: round(1e3 * float(np.mean(v)), 1) for k, v in sections.items()}, "ms per step")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
