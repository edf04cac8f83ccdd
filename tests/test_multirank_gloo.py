"""CPU suite: the N>1 bench logic (stream sharding, barrier, max-over-ranks timing, sum of per-rank work) with
world_size 2 over gloo.  No GPU and no decode here - the data path has no collective to test."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(n_streams_total, world, rank):
    """stream s lives on rank s mod world"""
    return [s for s in range(n_streams_total) if s % world == rank]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard(7, world, rank)
    samples = torch.tensor([float(len(mine) * 1000)], dtype=torch.float64)
    t = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)     # rank 1 is slower
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(samples, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((float(t), float(samples), mine))
    dist.destroy_process_group()


def test_two_ranks_shard_streams_and_reduce_timing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, samples, mine = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert abs(t - 0.020) < 1e-12          # max over ranks
    assert samples == 7000.0               # every stream counted exactly once
    assert mine == [0, 2, 4, 6]
    assert sorted(shard(7, 2, 0) + shard(7, 2, 1)) == list(range(7))


def test_stream_schedule_is_rank_independent():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.stream_bursts(5, 40_000_000)
    b = bench.stream_bursts(5, 40_000_000)
    assert [(x.at, x.sensor, x.frame) for x in a] == [(x.at, x.sensor, x.frame) for x in b] and len(a) >= 2
