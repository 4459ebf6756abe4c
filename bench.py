#!/usr/bin/env python
"""
bench.py — large-v3 RTFx (audio seconds / wall second) of the WhisperX hot path on B200.

  python bench.py --gpus 1 --steps K --warmup W                 our arm (CUDA kernels through the C-ABI)
  python bench.py --impl reference --gpus 1 --steps K --warmup W  the reference's CPU path (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...             one rank per GPU, weak scaling

One "step" = one pass of the hot path over one 30-minute batch of synthetic audio per GPU:
log-mel -> encoder -> batched greedy decode (all 224 sampled positions: random-init weights never emit
EOT) -> CTC trellis + beam-2 backtrack over synthetic wav2vec2-shaped emissions, 60 x 30 s VAD chunks.
`value` times the device-resident path; `e2e` times the public API (whisperx.load_model(...).transcribe
+ alignment from host emissions) with host buffers, H2D / D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "whisperx-mlx_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from fake_ctc_model import synthetic_speech  # noqa: E402  (deterministic synthetic audio generator)

SR = 16000
CHUNK_S = 30.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def workload(model, minutes, seed):
    n_chunks = int(round(minutes * 60 / CHUNK_S))
    base = synthetic_speech(60.0, seed=seed)  # 60 s of deterministic speech-like signal, tiled (generation is host-side prep)
    reps = int(np.ceil(n_chunks * CHUNK_S / 60.0))
    audio = np.tile(base, reps)[: int(n_chunks * CHUNK_S * SR)]
    rng = np.random.RandomState(seed)
    # alignment inputs: wav2vec2-shaped emissions T=1499, V=29; transcripts N~U(50,450), 5 % wildcards
    T, V = 1499, 29
    emis = (np.random.RandomState(seed + 1).standard_normal((n_chunks, T, V)) * 3.0).astype(np.float32)
    toks = []
    for _ in range(n_chunks):
        n = int(rng.randint(50, 451))
        t = rng.randint(1, V, size=n).astype(np.int32)
        t[rng.rand(n) < 0.05] = -1
        toks.append(t)
    return audio, n_chunks, emis, toks


def flops_encoder(dims):
    d, nm, L = dims["n_audio_state"], dims["n_mels"], dims["n_audio_layer"]
    return 2 * 3000 * d * 3 * nm + 2 * 1500 * d * 3 * d + L * (8 * 1500 * d * d + 4 * 1500 * 1500 * d + 16 * 1500 * d * d)


def decode_bytes_per_step(dims, B, t_mean):
    d, L, V = dims["n_text_state"], dims["n_text_layer"], dims["n_vocab"]
    weights = 2 * (L * 14 * d * d + V * d)       # once per step
    cross = B * 2 * (L * 2 * 1500 * d)
    selfkv = B * 2 * (L * 2 * t_mean * d)
    return weights, cross, selfkv


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the path (oracle port), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.pipeline import cpu_hot_path
    from whisperx.backends import b200_weights as bw
    torch.set_num_threads(os.cpu_count())
    dims = bw.dims_for(args.model)
    sp = bw.special_tokens(dims)
    w = bw.round_to_bf16(bw.init_random_weights(dims, seed=0))
    audio, n_chunks, emis, toks = workload(args.model, args.minutes, 1234)
    n = args.cpu_chunks
    chunks = [audio[i * 480000:(i + 1) * 480000] for i in range(n)]
    prompt = [sp["sot"], sp["sot"] + 1, sp["transcribe"], sp["no_timestamps"]]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, parts = cpu_hot_path(chunks, dims, w, prompt, sp["eot"], sp["no_speech"], dims["n_text_ctx"] // 2,
                                [emis[i] for i in range(n)], [toks[i].tolist() for i in range(n)])
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    value = n * CHUNK_S / sec
    sample = f"{n} of {n_chunks} chunks ({n * CHUNK_S:.0f} s audio), batch {n}, all 224 decode positions, mel+encoder+decoder+ctc"
    line = {"impl": "reference", "metric": "large-v3 RTFx (audio s / wall s)", "value": value, "unit": "x realtime",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"whisper-{args.model}, {args.minutes:g} min synthetic audio, 30 s VAD chunks (bounded sample per step)",
                       "parallelism": "cpu"},
            "cpu_baseline": {"value": value, "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                             "stage_seconds": parts},
            "e2e": {"value": value, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--minutes", type=float, default=30.0)
    ap.add_argument("--batch-size", type=int, default=60)
    ap.add_argument("--cpu-chunks", type=int, default=1)
    ap.add_argument("--sample-len", type=int, default=0, help="override the number of sampled positions (profiling only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from whisperx._native import CTC_BEAM2
    from whisperx.alignment import align_from_emissions
    import whisperx

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    import warnings
    warnings.simplefilter("ignore")
    pipe = whisperx.load_model(args.model, device="cuda", device_index=local, backend="b200", language="en",
                               vad_method="uniform", batch_size=args.batch_size)
    be = pipe.backend
    ctx, dims = be.ctx, be.dims
    audio, n_chunks, emis, toks = workload(args.model, args.minutes, 1234 + rank)  # every rank: its own 30 min (weak scaling)
    chunks = [audio[i * 480000:(i + 1) * 480000] for i in range(n_chunks)]
    audio_s = n_chunks * CHUNK_S

    # device-resident inputs for `value`
    audio_dev, offs, lens = be.upload_chunks(chunks)
    T, V = emis.shape[1], emis.shape[2]
    emis_dev = torch.from_numpy(emis.reshape(-1, V)).to(dev)
    emis_work = torch.empty_like(emis_dev)
    tok_dev = torch.from_numpy(np.concatenate(toks)).to(dev)
    t_off = (np.arange(n_chunks + 1) * T).astype(np.int32)
    n_off = np.concatenate([[0], np.cumsum([len(t) for t in toks])]).astype(np.int32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    prompt = be.tokenizer.prompt("en", "transcribe", True)
    if args.sample_len > 0:
        be.options["sample_len"] = args.sample_len
    sample_len = int(be.options["sample_len"])

    stage_ev = {k: [] for k in ("mel", "encoder", "decode", "ctc")}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step_device(record):
        flush.zero_()
        for i in range(0, n_chunks, args.batch_size):
            j = min(n_chunks, i + args.batch_size)
            e0 = ev()
            mel = ctx.logmel(audio_dev, offs[i:j], lens[i:j], 480000, dims["n_mels"], be._filters)
            e1 = ev()
            enc = ctx.encode(mel)
            e2 = ev()
            r = ctx.decode_greedy(enc, prompt, be.specials["eot"], no_speech=be.specials["no_speech"], sample_len=sample_len,
                                  suppress_blank=True, blank_token=be.specials["blank"])
            e3 = ev()
            if record:
                stage_ev["mel"].append((e0, e1)); stage_ev["encoder"].append((e1, e2)); stage_ev["decode"].append((e2, e3))
        e4 = ev()
        emis_work.copy_(emis_dev)
        ctx.log_softmax_rows_(emis_work)
        res = ctx.ctc_align(emis_work, t_off, tok_dev, n_off, 0, CTC_BEAM2)
        e5 = ev()
        if record:
            stage_ev["ctc"].append((e4, e5))
        return r, res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device(False)
    ctx.decode_stats(reset=True)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step_device(True)
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    launches = ctx.launches - launches0
    clock_info = clocks.stop() if rank == 0 else None
    cross_ms, steps_ms, n_dec_steps = ctx.decode_stats(reset=True)
    stage_ms = {k: float(sum(a.elapsed_time(b) for a, b in v)) / args.steps for k, v in stage_ev.items()}
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * audio_s / (ms_step / 1e3)

    # ---- e2e through the public API with host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        emis_list = [emis[i] for i in range(n_chunks)]
        tok_lists = [t.tolist() for t in toks]

        def step_e2e():
            out = pipe.transcribe(audio, batch_size=args.batch_size, chunk_size=30)  # numpy in -> dicts out (H2D + D2H inside)
            paths = align_from_emissions(emis_list, tok_lists, 0, device_index=local)  # pinned H2D -> K4 -> D2H
            return out, paths

        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out, paths = step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        sec = float(te.item()) / args.steps
        h2d = audio.nbytes + emis.nbytes + sum(t.nbytes for t in toks)
        d2h = n_chunks * (sample_len * 4 + 12) + n_chunks * T * 8 + n_chunks * 4
        e2e = {"value": world * audio_s / sec, "unit": "x realtime", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": sec * 1e3, "segments_returned": len(out["segments"]), "aligned_ok": int(sum(1 for p in paths if p[0] == 0))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: dec_step_kernel, the persistent cooperative decode kernel.  One launch
    # walks STEPS_PER_LAUNCH decode positions (wxb_decode_opts.check_every); its duration is measured with CUDA
    # events on the launching stream inside wxb_decode_greedy (wxb_decode_stats), summed over the timed region.
    P = peaks()
    B = args.batch_size
    prompt_len = len(prompt)
    t_mean = (prompt_len + sample_len) / 2.0
    wbytes, cbytes, sbytes = decode_bytes_per_step(dims, min(B, n_chunks), t_mean)
    bytes_step = wbytes + cbytes + sbytes  # algorithmic: weights once per step, cross-KV + self-KV per sequence
    steps_per_pass = n_dec_steps / args.steps
    dec_ms_per_step = steps_ms / max(n_dec_steps, 1)
    steps_per_launch = min(16, sample_len)
    achieved = bytes_step / (dec_ms_per_step * 1e-3) / 1e9
    enc_flops = flops_encoder(dims) * n_chunks
    ckv_flops = 2 * 1500 * dims["n_text_state"] * dims["n_audio_state"] * 2 * dims["n_text_layer"] * n_chunks
    mel_bytes = n_chunks * (4 * 480000 + dims["n_mels"] * 3000 * 4)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dec_step_traffic.json")  # from the committed `ncu --set full` capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("model") == be.model_name and tj.get("batch") == B:
            traffic = tj["dram_bytes_per_launch"]
    roofline = {"bound": "hbm", "kernel": "dec_step_kernel (persistent cooperative kernel: %d operators per decode step separated by grid barriers, "
                                          "%d steps per launch)" % (11 * dims["n_text_layer"] + 3, steps_per_launch),
                "achieved": achieved, "peak": P["hbm_gbs"], "unit": "GB/s", "frac": achieved / P["hbm_gbs"], "traffic": traffic,
                "peak_source": P["source"], "bytes_per_launch": bytes_step * steps_per_launch,
                "ms_per_launch": dec_ms_per_step * steps_per_launch, "steps_per_launch": steps_per_launch,
                "bytes_per_step": bytes_step, "ms_per_step": dec_ms_per_step,
                "stages": {
                    "mel": {"ms": stage_ms["mel"], "GB/s": mel_bytes / (stage_ms["mel"] * 1e-3) / 1e9, "frac_hbm": mel_bytes / (stage_ms["mel"] * 1e-3) / 1e9 / P["hbm_gbs"]},
                    "encoder": {"ms": stage_ms["encoder"], "TFLOP/s": enc_flops / (stage_ms["encoder"] * 1e-3) / 1e12,
                                "frac_bf16_burst": enc_flops / (stage_ms["encoder"] * 1e-3) / 1e12 / P["bf16_tflops"],
                                "frac_bf16_sustained": enc_flops / (stage_ms["encoder"] * 1e-3) / 1e12 / (P["bf16_tflops_sustained"] or P["bf16_tflops"])},
                    "cross_kv_gemm": {"ms": cross_ms / args.steps, "TFLOP/s": ckv_flops / (cross_ms / args.steps * 1e-3) / 1e12},
                    "decode_steps": {"ms": steps_ms / args.steps, "steps": steps_per_pass, "GB/s": achieved},
                    "ctc": {"ms": stage_ms["ctc"]}}}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        from oracle.pipeline import cpu_hot_path
        from whisperx.backends import b200_weights as bw
        torch.set_num_threads(os.cpu_count())
        w = bw.kernel_layout_to_openai_fp32(be.kernel_weights, dims)  # the very numbers the GPU used, as fp32
        n = args.cpu_chunks
        t0 = time.perf_counter()
        _, parts = cpu_hot_path(chunks[:n], dims, w, prompt, be.specials["eot"], be.specials["no_speech"], sample_len,
                                [emis[i] for i in range(n)], [toks[i].tolist() for i in range(n)])
        sec = time.perf_counter() - t0
        cpu_baseline = {"value": n * CHUNK_S / sec, "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} of {n_chunks} chunks ({n * CHUNK_S:.0f} s audio), batch {n}, all {sample_len} decode positions, "
                                  "mel+encoder+decoder+ctc (torch CPU fp32 stands in for faster-whisper/CTranslate2, see DESIGN.md)",
                        "seconds": sec, "stage_seconds": parts}

    line = {"metric": "large-v3 RTFx (audio s / wall s)", "value": value, "unit": "x realtime", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"whisper-{be.model_name} (random-init), {args.minutes:g} min synthetic audio per GPU = {n_chunks} x 30 s VAD chunks, "
                                   f"batch {B}, log-mel + encoder + greedy decode ({sample_len} positions) + CTC beam-2 alignment "
                                   f"(T=1499, V=29 synthetic emissions; wav2vec2 forward not in the timed path)",
                       "batch_size": B, "parallelism": f"dp{world} (chunk-sharded, no collective)", "l2": "256 MB flush buffer written before every step"},
            "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
