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
    dist.gather_object(local, bucket, dst=dst)
    if rank != dst:
        return None
    merged = [s for r in bucket for s in r["segments"]]
    merged.sort(key=lambda s: (s["start"], s["end"]))
    out = {"segments": merged, "language": bucket[0].get("language", "en")}
    if any("word_segments" in r for r in bucket):
        words = [w for r in bucket for w in r.get("word_segments", [])]
        out["word_segments"] = sorted(words, key=lambda w: (w.get("start", float("inf"))))
    return out
