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
