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

# ---- pass 1, WITHOUT a profiler (cProfile inflates the Python-heavy host parts several times over): wall-clock sums of a few
# coarse functions, wrapped in place
import whisperx.alignment as _al  # noqa: E402


def _wrap(obj, name, key):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            sections.setdefault(key, []).append(time.perf_counter() - t0)
    setattr(obj, name, g)
    return f


if os.environ.get("E2E_GROUPS"):  # A/B of align()'s pipeline depth: "groups,min_segments"
    _al.PIPELINE_GROUPS, _al.PIPELINE_MIN_SEGMENTS = (int(v) for v in os.environ["E2E_GROUPS"].split(","))
be = pipe.backend
saved = [(be, "upload_chunks", _wrap(be, "upload_chunks", "  transcribe: upload_chunks (pinned copy + H2D enqueue)")),
         (be, "transcribe_device", _wrap(be, "transcribe_device", "  transcribe: transcribe_device (mel + encoder + decode, blocks in the EOT polls)")),
         (be.tokenizer, "decode", _wrap(be.tokenizer, "decode", "  transcribe: tokenizer.decode (sum over segments)")),
         (be, "transcribe_batch", _wrap(be, "transcribe_batch", "transcribe_batch")),
         (_al, "_prepare_segment", _wrap(_al, "_prepare_segment", "  align: _prepare_segment (sum)")),
         (_al, "_assemble", _wrap(_al, "_assemble", "  align: _assemble (sum)")),
         (_al, "_merge_runs", _wrap(_al, "_merge_runs", "  align: _merge_runs (sum)")),
         (bundle[0], "emissions", _wrap(bundle[0], "emissions", "  align: model.emissions (upload + forward enqueue, sum over groups)"))]
for _ in range(steps):
    step()
n = steps
print("un-profiled pass, ms per step:")
for k, v in sections.items():
    print("  %-90s %8.1f" % (k, 1e3 * sum(v) / n))
for obj, name, f in saved:
    setattr(obj, name, f)
# timeline of one align() call (whisperx.alignment.TRACE)
_al.TRACE = []
step()
tr, _al.TRACE = _al.TRACE, None
print("align() timeline of one step, ms since its start:")
for label, t in tr:
    print("  %8.2f  %s" % (1e3 * (t - tr[0][1]), label))
sections.clear()
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    step()
pr.disable()
print({k: round(1e3 * float(np.mean(v)), 1) for k, v in sections.items()}, "ms per step")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
