"""
Multi-GPU sharding of the hot path: VAD chunks are independent, so they are dealt to
one queue per GPU (one process per GPU) with longest-processing-time-first; there is no data-path
collective — only a final host-side gather of the per-chunk result dicts.
"""
from typing import Dict, List, Sequence


def lpt_partition(durations: Sequence[float], n_parts: int) -> List[List[int]]:
    """Indices of `durations` dealt into n_parts queues, longest first onto the lightest queue."""
    order = sorted(range(len(durations)), key=lambda i: (-durations[i], i))
    loads = [0.0] * n_parts
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for i in order:
        k = min(range(n_parts), key=lambda q: (loads[q], q))
        parts[k].append(i)
        loads[k] += durations[i]
    for p in parts:
        p.sort()
    return parts


def shard_segments(segments: List[Dict], rank: int, world_size: int) -> List[Dict]:
    parts = lpt_partition([s["end"] - s["start"] for s in segments], world_size)
    return [segments[i] for i in parts[rank]]


def gather_results(local: Dict, rank: int, world_size: int, dst: int = 0):
    """Host gather of {"segments": [...], "language": ...} dicts onto rank `dst`, merged and sorted by
    start time.  Uses torch.distributed's object gather (gloo or nccl process group)."""
    import torch.distributed as dist
    if world_size == 1:
        return local
    bucket = [None] * world_size if rank == dst else None
    dist.gather_object(plain_numbers(local), bucket, dst=dst)
    if rank != dst:
        return None
    merged = [s for r in bucket for s in r["segments"]]
    merged.sort(key=lambda s: (s["start"], s["end"]))
    out = {"segments": merged, "language": bucket[0].get("language", "en")}
    if any("word_segments" in r for r in bucket):
        out["word_segments"] = _merge_words(bucket, merged)
    return out


def _merge_words(bucket, merged) -> List[Dict]:
    """"word_segments" of the gathered result.  align() builds it by concatenating the segments' "words" in segment order
    (alignment.py:375-378); when every rank's list is exactly that, the same concatenation over the merged (start-sorted)
    segments is the single-process answer and needs no sort of the words.  Lists built any other way are merged by start time."""
    if all(len(r.get("word_segments", [])) == sum(len(s.get("words") or ()) for s in r["segments"]) for r in bucket):
        return [w for s in merged for w in (s.get("words") or ())]
    words = [w for r in bucket for w in r.get("word_segments", [])]
    return sorted(words, key=lambda w: (w.get("start", float("inf"))))


def plain_numbers(result: Dict) -> Dict:
    """Word / char times and scores of an aligned result as plain Python floats, in place (align() returns numpy.float64 scalars
    like the reference's pandas aggregation does; equal values, but pickle spends ~10 us on every numpy scalar: 19 ms to
    serialise one rank's 8 segments, 3 ms per rank to load them on rank 0 - most of the multi-GPU e2e gap).  The dicts of
    "word_segments" are the same objects as the segments' "words", so one pass covers both."""
    for seg in result.get("segments", []):
        for key in ("words", "chars"):
            for w in seg.get(key) or ():
                for k in ("start", "end", "score"):
                    if k in w:
                        w[k] = float(w[k])
    return result


def host_group(world_size: int):
    """A gloo process group for the host-side gather (None when single-process): result dicts never touch the GPUs."""
    import torch.distributed as dist
    if world_size == 1:
        return None
    return dist.new_group(backend="gloo")


def transcribe_sharded(pipeline, audio, rank: int, world_size: int, batch_size: int = 8, chunk_size: int = 30,
                       group=None, align_fn=None, **kwargs):
    """ONE transcription job on `world_size` GPUs (one process per GPU, replicated weights): the VAD-cut chunks of `audio`
    are dealt to per-GPU queues (LPT), every rank transcribes its own queue in batches of `batch_size` on its own device,
    and the per-chunk result dicts are gathered on rank 0 and merged by start time.  No data-path collective; the only
    exchange is the final host gather (process pattern of /root/reference/whisperx/process_separation.py:87-172).
    `align_fn(local_result, local_segments)` optionally aligns on the rank that transcribed (segments stay on their GPU).
    Returns the merged result on rank 0, None elsewhere."""
    import torch.distributed as dist
    from .audio import SAMPLE_RATE
    if pipeline.vad_model is None:
        from .vads import synthetic_vad_cuts
        segments = synthetic_vad_cuts(len(audio) / SAMPLE_RATE, "uniform", chunk_size=chunk_size)
    else:
        segments = pipeline._segment_audio_with_vad(audio, chunk_size)
    mine = shard_segments(segments, rank, world_size)
    for seg in mine:
        seg["audio"] = audio[int(seg["start"] * SAMPLE_RATE): int(seg["end"] * SAMPLE_RATE)]
    local = pipeline.backend.transcribe_batch(mine, batch_size=batch_size, **kwargs)
    if align_fn is not None:
        local = align_fn(local, mine)
    if world_size == 1:
        return local
    bucket = [None] * world_size if rank == 0 else None
    dist.gather_object(plain_numbers(local), bucket, dst=0, group=group)
    if rank != 0:
        return None
    merged = [s for r in bucket for s in r["segments"]]
    merged.sort(key=lambda s: (s["start"], s["end"]))
    out = {"segments": merged, "language": bucket[0].get("language", "en")}
    if any("word_segments" in r for r in bucket):
        out["word_segments"] = _merge_words(bucket, merged)
    return out
