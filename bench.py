#!/usr/bin/env python
"""
bench.py — large-v3 RTFx (audio seconds / wall second) of the WhisperX hot path on B200.

  python bench.py --gpus 1 --steps K --warmup W                   our arm (CUDA kernels through the C-ABI)
  python bench.py --impl reference --gpus 1 --steps K --warmup W  the reference's CPU path (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...               one rank per GPU

One "step" = one pass of the hot path over ONE 30-minute job of synthetic audio: log-mel -> encoder -> batched greedy
decode (all 224 sampled positions: random-init weights never emit EOT) -> wav2vec2 emissions -> CTC trellis + beam-2
backtrack, 60 x 30 s VAD chunks.  With N > 1 the job's chunks are dealt to the N GPUs (LPT queues, whisperx/multi_gpu.py),
every rank works through its own queue and the per-chunk results are gathered on rank 0 (host gather, no data-path
collective): `value` is STRONG scaling of one job.  The weak-scaling figure (every rank its own 30 minutes) is the sub-field
`weak`.  `value` times the device-resident path; `e2e` times the public API (whisperx.load_model(...).transcribe... +
whisperx.align()) with host buffers, H2D / D2H and the host gather inside the timed region.

Sub-records in the same JSON line: `batch8` (BASELINE config 3: large-v3, 8 chunks, batch 8, N = 1 only) and `turbo`
(config 4: large-v3-turbo, the same 30-minute job sharded like the headline).
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
# environment knobs of the library: tracing ones are echoed, anything else aborts the run (a bench number must come from the
# default code path of the in-tree library)
ENV_TRACING = {"WXB_DEC_PROF"}


def env_guard(allow):
    found = {k: v for k, v in os.environ.items() if k.startswith("WXB")}
    bad = [k for k in found if k not in ENV_TRACING and not allow]
    if bad:
        sys.exit(f"bench.py: refusing to run with {bad} set (library variant / probe knobs); unset them or pass --allow-env "
                 "for an A/B run whose number is not a bench value")
    return found


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


def job_audio(minutes, seed):
    """The job: `minutes` of deterministic speech-like audio (60 s pattern tiled; generation is host-side preparation)."""
    n = int(round(minutes * 60 * SR))
    base = synthetic_speech(60.0, seed=seed)
    return np.tile(base, int(np.ceil(n / len(base))))[:n]


def align_inputs(n_chunks, seed):
    """Alignment inputs of SURVEY 8d for callers WITHOUT an alignment model (the CPU arm): wav2vec2-shaped emissions
    T = 1499, V = 29 and transcripts N ~ U(50, 450) with 5 % wildcards."""
    T, V = 1499, 29
    rng = np.random.RandomState(seed)
    emis = (np.random.RandomState(seed + 1).standard_normal((n_chunks, T, V)) * 3.0).astype(np.float32)
    toks = []
    for _ in range(n_chunks):
        n = int(rng.randint(50, 451))
        t = rng.randint(1, V, size=n).astype(np.int32)
        t[rng.rand(n) < 0.05] = -1
        toks.append(t)
    return emis, toks


def flops_encoder(dims):
    d, nm, L = dims["n_audio_state"], dims["n_mels"], dims["n_audio_layer"]
    return 2 * 3000 * d * 3 * nm + 2 * 1500 * d * 3 * d + L * (8 * 1500 * d * d + 4 * 1500 * 1500 * d + 16 * 1500 * d * d)


def decode_bytes_per_step(dims, B, t_mean):
    d, L, V = dims["n_text_state"], dims["n_text_layer"], dims["n_vocab"]
    weights = 2 * (L * 14 * d * d + V * d)       # once per step
    cross = B * 2 * (L * 2 * 1500 * d)
    selfkv = B * 2 * (L * 2 * t_mean * d)
    return weights, cross, selfkv


def workload_config(model_name, minutes, n_chunks, world, n_mine, batch_size, sample_len, align_note):
    """The job both arms are measured on, in words (the reference arm times a bounded sample of it on the host cores)."""
    return {"workload": f"whisper-{model_name} (random-init), ONE {minutes:g} min synthetic recording = {n_chunks} x 30 s VAD chunks "
                        f"dealt to {world} GPU(s) (LPT queues, {n_mine} chunks on rank 0), batch {min(batch_size, n_mine)}, log-mel + "
                        f"encoder + greedy decode ({sample_len} positions) + {align_note}",
            "batch_size": batch_size, "parallelism": f"{world} x 1 GPU, chunk-sharded, host gather only (no collective)"}


ALIGN_NOTE = "wav2vec2-base forward (own kernels, random-init) + CTC beam-2 alignment"


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
    audio = job_audio(args.minutes, 1234)
    n_chunks = int(round(args.minutes * 60 / CHUNK_S))
    emis, toks = align_inputs(n_chunks, 1234)
    n = args.cpu_chunks
    chunks = [audio[i * 480000:(i + 1) * 480000] for i in range(n)]
    world = max(1, int(os.environ.get("WORLD_SIZE", str(args.gpus))))
    n_mine = -(-n_chunks // world)  # the LPT deal of equal chunks gives rank 0 the ceiling
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {**workload_config(args.model, args.minutes, n_chunks, world, n_mine, args.batch_size, dims["n_text_ctx"] // 2, ALIGN_NOTE),
                       "arm": "the same job's hot path on the host cores (oracle port), a bounded sample per step: " + sample +
                              "; the wav2vec2 forward is not in the CPU sample"},
            "cpu_baseline": {"value": value, "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                             "stage_seconds": parts},
            "e2e": {"value": value, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
class Job:
    """One model on this rank's GPU + this rank's share of a job's chunks, resident in HBM."""

    def __init__(self, args, model, local, rank, world, audio, segments, shard, batch_size, align_bundle):
        import whisperx
        from whisperx.multi_gpu import shard_segments
        self.args, self.rank, self.world, self.local = args, rank, world, local
        self.pipe = whisperx.load_model(model, device="cuda", device_index=local, backend="b200", language="en",
                                        vad_method="uniform", batch_size=batch_size, align_model=align_bundle)
        self.be = self.pipe.backend
        self.ctx, self.dims = self.be.ctx, self.be.dims
        if args.dec_groups:            # A/B aids of the decoder's sequence-group schedule (results do not depend on them)
            self.ctx.debug_set("dec_groups", args.dec_groups)
        if args.dec_group_delay_ns >= 0:
            self.ctx.debug_set("dec_group_delay_ns", args.dec_group_delay_ns)
        if args.enc_group >= 0:        # A/B aid: chunks per group of the encoder's layer stack (results do not depend on it)
            self.ctx.debug_set("enc_group", args.enc_group)
        if args.sample_len > 0:
            self.be.options["sample_len"] = args.sample_len
        self.sample_len = int(self.be.options["sample_len"])
        self.batch_size = batch_size
        self.audio = audio
        self.segments = segments
        self.mine = shard_segments(segments, rank, world) if shard else list(segments)
        self.chunks = [audio[int(s["start"] * SR): int(s["end"] * SR)] for s in self.mine]
        self.n_mine = len(self.chunks)
        self.audio_dev, self.offs, self.lens = self.be.upload_chunks(self.chunks) if self.chunks else (None, [], [])
        self.prompt = self.be.tokenizer.prompt("en", "transcribe", True)
        self.align_bundle = align_bundle
        self.stage_ev = {k: [] for k in ("mel", "encoder", "decode", "w2v", "ctc")}

    def step_device(self, record):
        """Device-resident pass over this rank's chunks: K1 -> K2 -> K3, then wav2vec2 emissions + K4 for every chunk."""
        ctx, be, dims = self.ctx, self.be, self.dims

        def ev():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            return e

        be._bind()
        toks_dev = []
        for i in range(0, self.n_mine, self.batch_size):
            j = min(self.n_mine, i + self.batch_size)
            e0 = ev()
            ctx.logmel_features(self.audio_dev, self.offs[i:j], self.lens[i:j], dims["n_mels"], be._filters)  # K1 -> K2 on the device
            e1 = ev()
            enc = ctx.encode(None, n_chunks=j - i)
            e2 = ev()
            r = ctx.decode_greedy(enc, self.prompt, be.specials["eot"], no_speech=be.specials["no_speech"], sample_len=self.sample_len,
                                  suppress_blank=True, blank_token=be.specials["blank"])
            e3 = ev()
            toks_dev.append(r["tokens"])
            if record:
                self.stage_ev["mel"].append((e0, e1)); self.stage_ev["encoder"].append((e1, e2)); self.stage_ev["decode"].append((e2, e3))
        if self.align_bundle is not None and self.n_mine:
            self.align_device(record, ev)
        return toks_dev

    def align_device(self, record, ev):
        """Alignment leg with everything resident: wav2vec2 forward over this rank's chunks (own kernels) + log_softmax + K4 on
        fixed synthetic transcripts (the decoded text needs the host tokenizer; that round trip is what `e2e` times)."""
        from whisperx._native import CTC_BEAM2
        model, _meta = self.align_bundle
        e4 = ev()
        emis, t_off = model.emissions_device(self.audio_dev, self.offs, self.lens)  # [sumT, V] f32 logits, frame offsets
        e5 = ev()
        self.ctx.log_softmax_rows_(emis)
        if getattr(self, "_ctc_tok", None) is None:
            rng = np.random.RandomState(4321 + self.rank)
            V = emis.shape[1]
            toks = []
            for k in range(self.n_mine):
                T = int(t_off[k + 1] - t_off[k])
                n = int(rng.randint(50, max(51, min(451, T))))
                t = rng.randint(1, V, size=n).astype(np.int32)
                t[rng.rand(n) < 0.05] = -1
                toks.append(t)
            self._ctc_tok = torch.from_numpy(np.concatenate(toks)).to(self.ctx.device)
            self._ctc_noff = np.concatenate([[0], np.cumsum([len(t) for t in toks])]).astype(np.int32)
        res = self.ctx.ctc_align(emis, t_off, self._ctc_tok, self._ctc_noff, 0, CTC_BEAM2)
        e6 = ev()
        if record:
            self.stage_ev["w2v"].append((e4, e5)); self.stage_ev["ctc"].append((e5, e6))
        return res


def timed(job, dist, world, steps, warmup, flush, sampler=None):
    """W warm-up passes, then `steps` timed passes bracketed by barrier + synchronize; returns (ms per step = max over ranks,
    launches, decode stats, stage ms)."""
    dev = job.ctx.device

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        job.step_device(False)
    job.ctx.decode_stats(reset=True)
    for v in job.stage_ev.values():
        v.clear()
    barrier()
    if sampler is not None:
        sampler.start()
    launches0 = job.ctx.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        flush.zero_()  # > L2: nothing of the previous pass is cache-resident
        job.step_device(True)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clock_info = sampler.stop() if sampler is not None else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cross_ms, steps_ms, n_dec = job.ctx.decode_stats(reset=True)
    stage_ms = {k: float(sum(a.elapsed_time(b) for a, b in v)) / steps for k, v in job.stage_ev.items()}
    return dict(ms_step=float(t.item()) / steps, launches=job.ctx.launches - launches0, cross_ms=cross_ms / steps,
                dec_ms=steps_ms / steps, n_dec=n_dec / steps, stage_ms=stage_ms, clocks=clock_info)


def side_paths(job, audio, flush, steps):
    """Device time of the two side paths of SURVEY 8(f) on the headline job, CUDA events on the launching stream:
    dtw_words  the decode with the alignment heads' queries logged, then scores + cost rows + DTW paths for every sequence
               (224 token rows x 1500 frames each), and the host grouping of the words;
    vad        log-energy frame scores of the 30-minute recording + Binarize / merge_chunks in one kernel."""
    from whisperx.word_timing import dtw_word_timestamps
    ctx, be = job.ctx, job.be

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    n = min(job.n_mine, job.batch_size)
    heads = be._alignment_heads()
    ctx.logmel_features(job.audio_dev, job.offs[:n], job.lens[:n], job.dims["n_mels"], be._filters)
    enc = ctx.encode(None, n_chunks=n)
    dec_ms, dtw_ms, host_ms, n_words = [], [], [], 0
    for it in range(steps + 1):
        flush.zero_()
        ctx.collect_alignment_heads(heads)
        e0 = ev()
        r = ctx.decode_greedy(enc, job.prompt, be.specials["eot"], no_speech=be.specials["no_speech"], sample_len=job.sample_len,
                              suppress_blank=True, blank_token=be.specials["blank"])
        e1 = ev()
        n_rows = np.full(n, job.sample_len, dtype=np.int32)
        qk = ctx.dtw_scores(n_rows, len(job.prompt) - 1)
        cost = ctx.dtw_cost(qk)
        e2 = ev()
        paths = ctx.dtw_path(cost, n_rows)   # includes the D2H of the paths
        e3 = ev()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        toks = r["tokens"].cpu().numpy()
        words = dtw_word_timestamps(ctx, [toks[k] for k in range(n)], be.specials["eot"], len(job.prompt), be.tokenizer.decode_piece)
        t1 = time.perf_counter()
        if it:  # first pass = warm-up
            dec_ms.append(e0.elapsed_time(e1)); dtw_ms.append(e1.elapsed_time(e3)); host_ms.append((t1 - t0) * 1e3)
            n_words = sum(len(w) for w in words)
    ctx.collect_alignment_heads(None)
    dtw = {"config": f"{n} sequences x {job.sample_len} token rows x 1500 frames, {len(heads)} alignment heads ({be.model_name})",
           "decode_with_query_log_ms": float(np.mean(dec_ms)), "scores_cost_path_ms": float(np.mean(dtw_ms)),
           "through_host_api_ms": float(np.mean(host_ms)), "words": n_words,
           "note": "through_host_api_ms = word_timing.dtw_word_timestamps (3 launches + path D2H + word grouping on the host)"}
    wav = torch.from_numpy(audio).to(ctx.device)
    vad_ms = []
    for it in range(steps + 1):
        flush.zero_()
        e0 = ev()
        sc = ctx.vad_energy_scores(wav)
        (res,) = ctx.vad_chunks(sc, np.array([0, sc.numel()]), np.array([wav.numel()]), 30.0, onset=0.5, offset=0.363,
                                frame_duration=0.025, frame_step=0.010)
        e1 = ev()
        torch.cuda.synchronize()
        if it:
            vad_ms.append(e0.elapsed_time(e1))
    vad = {"config": f"{len(audio) / SR / 60:g} min recording: log-energy frame scores ({sc.numel()} frames) + Binarize / merge_chunks kernel",
           "ms": float(np.mean(vad_ms)), "chunks": int(len(res["chunks"])), "regions": int(len(res["regions"])),
           "GB/s_audio_read": wav.numel() * 4 / (float(np.mean(vad_ms)) * 1e-3) / 1e9}
    return dtw, vad


def decode_roofline(job, r, P, n_rows):
    """Algorithmic bytes of the decode steps this rank ran / their device time."""
    prompt_len = len(job.prompt)
    t_mean = (prompt_len + job.sample_len) / 2.0
    n_groups = max(1, -(-n_rows // job.batch_size))
    rows = n_rows / n_groups  # sequences per decode call (mean)
    wbytes, cbytes, sbytes = decode_bytes_per_step(job.dims, rows, t_mean)
    bytes_step = wbytes + cbytes + sbytes
    ms_per_step = r["dec_ms"] / max(r["n_dec"], 1)
    achieved = bytes_step / (ms_per_step * 1e-3) / 1e9
    return dict(bytes_per_step=bytes_step, ms_per_step=ms_per_step, achieved=achieved, frac=achieved / P["hbm_gbs"],
                rows_per_call=rows, steps=r["n_dec"])


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
    ap.add_argument("--no-extras", action="store_true", help="skip the batch8 / ragged / turbo / weak sub-records")
    ap.add_argument("--no-align", action="store_true", help="leave the wav2vec2 + CTC alignment leg out (profiling only)")
    ap.add_argument("--dec-groups", type=int, default=0, help="A/B aid: sequence groups per decode call (0 = the library's choice)")
    ap.add_argument("--enc-group", type=int, default=-1, help="A/B aid: chunks per group of the encoder's layer stack (-1 = the library's default, 0 = whole batch)")
    ap.add_argument("--dec-group-delay-ns", type=int, default=-1, help="A/B aid: start offset between the sequence groups (-1 = default)")
    ap.add_argument("--allow-env", action="store_true", help="run although WXB_* variables are set (A/B runs, not a bench value)")
    args = ap.parse_args()
    env_seen = env_guard(args.allow_env)
    if args.impl == "reference":
        return run_reference(args)

    import warnings
    import torch.distributed as dist
    import whisperx
    from whisperx import multi_gpu
    from whisperx._native import load_library
    from whisperx.vads import synthetic_vad_cuts
    warnings.simplefilter("ignore")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    hgroup = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        hgroup = multi_gpu.host_group(world)
    dev = torch.device("cuda", local)
    lib_path = load_library()._wxb_path

    # the job: ONE 30-minute recording, 60 uniform 30 s VAD chunks; identical on every rank
    audio = job_audio(args.minutes, 1234)
    audio_s = len(audio) / SR
    segments = synthetic_vad_cuts(audio_s, "uniform", chunk_size=CHUNK_S)
    n_chunks = len(segments)
    align_bundle = None
    if not args.no_align:
        align_bundle = whisperx.load_align_model("en", dev, model_name="WAV2VEC2_ASR_BASE_960H", random_init=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    job = Job(args, args.model, local, rank, world, audio, segments, True, args.batch_size, align_bundle)
    be, dims = job.be, job.dims
    sampler = ClockSampler(local) if rank == 0 else None
    r = timed(job, dist, world, args.steps, args.warmup, flush, sampler)
    value = audio_s / (r["ms_step"] / 1e3)
    w2v_stats = dict(getattr(align_bundle[0], "last_stats", {})) if align_bundle is not None else {}  # of THIS job (later jobs overwrite them)

    # ---- e2e: the public API with host buffers; chunks sharded, results gathered on rank 0 inside the timed region -------
    e2e = None
    if not args.no_e2e:
        def align_fn(local_result, mine):
            if align_bundle is None or not local_result["segments"]:
                return local_result
            model, meta = align_bundle
            out = whisperx.align(local_result["segments"], model, meta, audio, str(dev))
            out["language"] = local_result["language"]
            return out

        def step_e2e():
            return multi_gpu.transcribe_sharded(job.pipe, audio, rank, world, batch_size=args.batch_size, chunk_size=30,
                                                group=hgroup, align_fn=align_fn)

        for _ in range(min(args.warmup, 2)):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        sec = float(te.item()) / args.steps
        h2d = sum(int(c.nbytes) for c in job.chunks)
        d2h = job.n_mine * (job.sample_len * 4 + 12)
        if align_bundle is not None:
            st = getattr(align_bundle[0], "last_stats", {})
            h2d += int(st.get("h2d_bytes", 0))
            d2h += int(st.get("d2h_bytes", 0))
        tb = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        e2e = {"value": audio_s / sec, "unit": "x realtime", "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()),
               "ms_per_step": sec * 1e3}
        if rank == 0:
            e2e["segments_returned"] = len(out["segments"])
            e2e["words_returned"] = len(out.get("word_segments", []))
            e2e["api"] = ("whisperx.load_model(..., backend='b200') -> multi_gpu.transcribe_sharded (VAD cuts, LPT shard, "
                          "backend.transcribe_batch, whisperx.align on the transcribing rank, host gather on rank 0)")

    # ---- sub-records ---------------------------------------------------------------------------------------------------------
    P = peaks()
    extras = {}
    if not args.no_extras:
        if world > 1:
            # weak scaling: every rank its own 30 minutes (what round 1 reported); one warm-up, `steps` passes
            wjob = Job(args, args.model, local, rank, world, job_audio(args.minutes, 1234 + rank), segments, False, args.batch_size,
                       align_bundle)
            wr = timed(wjob, dist, world, args.steps, 1, flush)
            extras["weak"] = {"value": world * audio_s / (wr["ms_step"] / 1e3), "unit": "x realtime", "ms_per_step": wr["ms_step"],
                              "audio_s_per_gpu": audio_s, "note": "every rank transcribes its own 30 min (no sharing, no gather)"}
            del wjob
        else:
            extras["weak"] = {"value": value, "unit": "x realtime", "ms_per_step": r["ms_step"], "audio_s_per_gpu": audio_s}
            # BASELINE config 3: large-v3, batch 8 (8 chunks of the same job)
            b8 = Job(args, args.model, local, 0, 1, audio, segments[:8], False, 8, None)
            r8 = timed(b8, dist, 1, args.steps, 2, flush)
            d8 = decode_roofline(b8, r8, P, 8)
            extras["batch8"] = {"config": "whisper-large-v3, 8 x 30 s chunks, batch 8 (BASELINE config 3), mel + encoder + greedy decode",
                                "value": 8 * CHUNK_S / (r8["ms_step"] / 1e3), "unit": "x realtime", "ms_per_step": r8["ms_step"],
                                "decode": {"ms_per_step": d8["ms_per_step"], "GB/s": d8["achieved"], "frac_hbm": d8["frac"],
                                           "bytes_per_step": d8["bytes_per_step"]},
                                "encoder_ms": r8["stage_ms"]["encoder"],
                                "encoder_TFLOP/s": flops_encoder(dims) * 8 / (r8["stage_ms"]["encoder"] * 1e-3) / 1e12}
            del b8
            # SURVEY 8 f-2 / f-3 on the headline job (N = 1): word timing from the decoder's cross-attention, VAD chunking
            extras["dtw_words"], extras["vad"] = side_paths(job, audio, flush, args.steps)
        # the same recording cut into ragged VAD chunks, durations ~ U(5, 30) s (about 100 chunks), LPT-dealt to the
        # ranks; every rank walks its queue in batches of `batch_size` (chunks are padded to 30 s by the log-mel, as in the reference)
        rseg = synthetic_vad_cuts(audio_s, "ragged", chunk_size=CHUNK_S)
        rj = Job(args, args.model, local, rank, world, audio, rseg, True, args.batch_size, align_bundle)
        rr = timed(rj, dist, world, max(1, min(args.steps, 2)), 1, flush)
        drr = decode_roofline(rj, rr, P, rj.n_mine)
        extras["ragged"] = {"config": f"whisper-{args.model}, the same {args.minutes:g} min cut into {len(rseg)} ragged VAD chunks (U(5, 30) s, seed 1234) "
                                      f"sharded over {world} GPU(s), batches of {args.batch_size}; {rj.n_mine} chunks on rank 0",
                            "value": audio_s / (rr["ms_step"] / 1e3), "unit": "x realtime", "ms_per_step": rr["ms_step"],
                            "decode_rank0": {"ms_per_step": drr["ms_per_step"], "GB/s": drr["achieved"], "frac_hbm": drr["frac"],
                                             "rows_per_call": drr["rows_per_call"]},
                            "note": "random-init weights never emit EOT, so every chunk decodes all sampled positions whatever its duration"}
        del rj
        # BASELINE config 4: large-v3-turbo, the same 30-minute job sharded over the ranks
        tj = Job(args, "large-v3-turbo", local, rank, world, audio, segments, True, args.batch_size, None)
        tr = timed(tj, dist, world, args.steps, 2, flush)
        dt_ = decode_roofline(tj, tr, P, tj.n_mine)
        extras["turbo"] = {"config": f"whisper-large-v3-turbo (4-layer decoder), {args.minutes:g} min = {n_chunks} chunks sharded over "
                                     f"{world} GPU(s) (BASELINE config 4), mel + encoder + greedy decode",
                           "value": audio_s / (tr["ms_step"] / 1e3), "unit": "x realtime", "ms_per_step": tr["ms_step"],
                           "decode_rank0": {"ms_per_step": dt_["ms_per_step"], "GB/s": dt_["achieved"], "frac_hbm": dt_["frac"],
                                            "rows_per_call": dt_["rows_per_call"]}}
        del tj
        be._bind()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: dec_step_kernel, the persistent cooperative decode kernel.  One launch walks up
    # to 16 decode positions (wxb_decode_opts.check_every); its duration is measured with CUDA events on the launching
    # stream inside wxb_decode_greedy (wxb_decode_stats), summed over the timed region (rank 0's share of the job).
    dr = decode_roofline(job, r, P, job.n_mine)
    steps_per_launch = min(16, job.sample_len)
    n_mine = job.n_mine
    enc_flops = flops_encoder(dims) * n_mine
    ckv_flops = 2 * 1500 * dims["n_text_state"] * dims["n_audio_state"] * 2 * dims["n_text_layer"] * n_mine
    mel_bytes = n_mine * (4 * 480000 + dims["n_mels"] * 3000 * 2)  # f32 samples in, bf16 frame-major features out (what the encoder reads)
    sm = r["stage_ms"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dec_step_traffic.json")  # from the committed `ncu --set full` capture
    if os.path.exists(tpath):
        tj_ = json.load(open(tpath))
        if tj_.get("model") == be.model_name and tj_.get("batch") == min(args.batch_size, n_mine):
            traffic = tj_["dram_bytes_per_launch"]
    stages = {
        "mel": {"ms": sm["mel"], "GB/s": mel_bytes / (sm["mel"] * 1e-3) / 1e9, "frac_hbm": mel_bytes / (sm["mel"] * 1e-3) / 1e9 / P["hbm_gbs"]},
        "encoder": {"ms": sm["encoder"], "TFLOP/s": enc_flops / (sm["encoder"] * 1e-3) / 1e12,
                    "frac_bf16_burst": enc_flops / (sm["encoder"] * 1e-3) / 1e12 / P["bf16_tflops"],
                    "frac_bf16_sustained": enc_flops / (sm["encoder"] * 1e-3) / 1e12 / (P["bf16_tflops_sustained"] or P["bf16_tflops"])},
        "cross_kv_gemm": {"ms": r["cross_ms"], "TFLOP/s": ckv_flops / (r["cross_ms"] * 1e-3) / 1e12},
        "decode_steps": {"ms": r["dec_ms"], "steps": dr["steps"], "GB/s": dr["achieved"]}}
    if align_bundle is not None:
        st = w2v_stats
        stages["wav2vec2"] = {"ms": sm["w2v"], "TFLOP/s": st.get("flops", 0) / max(sm["w2v"] * 1e-3, 1e-9) / 1e12, "frames": st.get("frames")}
        stages["ctc"] = {"ms": sm["ctc"]}
    roofline = {"bound": "hbm", "kernel": "dec_step_kernel (persistent cooperative kernel: %d operators per decode step separated by grid "
                                          "barriers, %d steps per launch)" % (11 * dims["n_text_layer"] + 3, steps_per_launch),
                "achieved": dr["achieved"], "peak": P["hbm_gbs"], "unit": "GB/s", "frac": dr["frac"], "traffic": traffic,
                "peak_source": P["source"], "bytes_per_launch": dr["bytes_per_step"] * steps_per_launch,
                "ms_per_launch": dr["ms_per_step"] * steps_per_launch, "steps_per_launch": steps_per_launch,
                "bytes_per_step": dr["bytes_per_step"], "ms_per_step": dr["ms_per_step"], "rows_per_call": dr["rows_per_call"],
                "stages": stages}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        from oracle.pipeline import cpu_hot_path
        from whisperx.backends import b200_weights as bw
        torch.set_num_threads(os.cpu_count())
        w = bw.kernel_layout_to_openai_fp32(be.kernel_weights, dims)  # the very numbers the GPU used, as fp32
        n = args.cpu_chunks
        emis, toks = align_inputs(n_chunks, 1234)
        t0 = time.perf_counter()
        _, parts = cpu_hot_path(job.chunks[:n], dims, w, job.prompt, be.specials["eot"], be.specials["no_speech"], job.sample_len,
                                [emis[i] for i in range(n)], [toks[i].tolist() for i in range(n)])
        sec = time.perf_counter() - t0
        cpu_baseline = {"value": n * CHUNK_S / sec, "unit": "x realtime", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} of {n_chunks} chunks ({n * CHUNK_S:.0f} s audio), batch {n}, all {job.sample_len} decode positions, "
                                  "mel+encoder+decoder+ctc (torch CPU fp32 stands in for faster-whisper/CTranslate2, see DESIGN.md; "
                                  "the wav2vec2 forward is not in the CPU sample)",
                        "seconds": sec, "stage_seconds": parts}

    align_note = (ALIGN_NOTE if align_bundle is not None
                  else "no alignment leg (--no-align)")
    line = {"metric": "large-v3 RTFx (audio s / wall s)", "value": value, "unit": "x realtime", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {**workload_config(be.model_name, args.minutes, n_chunks, world, n_mine, args.batch_size, job.sample_len, align_note),
                       "l2": "256 MB flush buffer written before every step", "library": lib_path, "env": env_seen,
                       **({"dec_groups": args.dec_groups} if args.dec_groups else {}),
                       **({"dec_group_delay_ns": args.dec_group_delay_ns} if args.dec_group_delay_ns >= 0 else {}),
                       **({"enc_group": args.enc_group} if args.enc_group >= 0 else {})},
            "clocks": r["clocks"], "e2e": e2e, "gpu_launches": int(r["launches"]), "roofline": roofline, "cpu_baseline": cpu_baseline}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
