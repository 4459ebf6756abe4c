"""N>1 host path on CPU: LPT sharding + object gather over a world_size-2 gloo group."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_lpt_partition_balances():
    from whisperx.multi_gpu import lpt_partition
    durs = [30, 30, 5, 7, 29, 12, 30, 1, 18, 22, 9, 30]
    parts = lpt_partition(durs, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(durs)))
    loads = [sum(durs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(durs)
    assert lpt_partition(durs, 1) == [list(range(len(durs)))]
    assert [len(p) for p in lpt_partition([30.0] * 60, 8)] == [8, 8, 8, 8, 7, 7, 7, 7]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from whisperx.multi_gpu import gather_results, shard_segments
    segs = [{"start": 30.0 * i, "end": 30.0 * i + (30.0 if i % 3 else 11.0)} for i in range(9)]
    mine = shard_segments(segs, rank, world)
    local = {"segments": [{"start": s["start"], "end": s["end"], "text": f"r{rank}"} for s in mine], "language": "en"}
    out = gather_results(local, rank, world)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the only collective the bench uses: timing max / counts
    if rank == 0:
        q.put((out, float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2_gloo():
    import sys
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 9.0
    starts = [s["start"] for s in out["segments"]]
    assert starts == sorted(starts) and len(starts) == 9
    assert {s["text"] for s in out["segments"]} == {"r0", "r1"}


class _FakeBackend:
    def __init__(self, rank):
        self.rank = rank

    def transcribe_batch(self, segments, batch_size=8, **kw):
        return {"segments": [{"start": s["start"], "end": s["end"], "text": f"r{self.rank}:{len(s['audio'])}"} for s in segments],
                "language": "en"}


class _FakePipeline:
    vad_model = "uniform"

    def __init__(self, rank):
        self.backend = _FakeBackend(rank)

    def _segment_audio_with_vad(self, audio, chunk_size):
        from whisperx.vads import synthetic_vad_cuts
        return synthetic_vad_cuts(len(audio) / 16000, "ragged", seed=5, chunk_size=chunk_size)


def _worker_sharded(rank, world, port, q):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from whisperx import multi_gpu
    group = multi_gpu.host_group(world)
    audio = np.zeros(16000 * 200, np.float32)
    out = multi_gpu.transcribe_sharded(_FakePipeline(rank), audio, rank, world, batch_size=4, chunk_size=30, group=group,
                                       align_fn=lambda res, mine: dict(res, word_segments=[{"word": "w", "start": s["start"]} for s in res["segments"]]))
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_transcribe_sharded_world2_gloo():
    """One job on two ranks: every VAD chunk is transcribed exactly once (by the rank LPT dealt it to), the merged result on
    rank 0 is sorted by start time and carries both ranks' segments and words."""
    from whisperx.vads import synthetic_vad_cuts
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cuts = synthetic_vad_cuts(200.0, "ragged", seed=5, chunk_size=30)
    assert [s["start"] for s in out["segments"]] == [c["start"] for c in cuts]
    assert [s["end"] for s in out["segments"]] == [c["end"] for c in cuts]
    ranks = {s["text"].split(":")[0] for s in out["segments"]}
    assert ranks == {"r0", "r1"}
    for s, c in zip(out["segments"], cuts):  # every rank saw the samples of its own chunks
        assert int(s["text"].split(":")[1]) == int(c["end"] * 16000) - int(c["start"] * 16000)
    assert len(out["word_segments"]) == len(cuts) and out["language"] == "en"
