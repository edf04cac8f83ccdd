"""CPU suite: the N>1 logic with world_size 2 over gloo - stream sharding, barrier, max-over-ranks timing, sum of
per-rank work (bench.py), and the two ways work is split over ranks, with the oracle standing in for the GPU so that
what is gathered can be compared with a single-process run: independent streams decoded on the rank that owns them,
and one buffer time-sharded with a lead-in for the decimation sweep (tools/sweep.py mgpu)."""
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


def _decode_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import iqsynth as g
    import oracle_lib as ol
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # (1) five independent streams, stream s on rank s mod world: every rank decodes its own, rank 0 gathers the -e lines
    mine = {}
    for s in shard(5, world, rank):
        o = ol.Oracle(types=0x07, thresh=0)
        o.process(g.fixture_single_tfa1(seed=1 + s))
        mine[s] = [r["exec"] for r in o.records()]
        o.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # (2) one buffer, time-sharded: slice + a lead-in of 64 outputs, which are dropped (tools/sweep.py mgpu)
    rng = np.random.default_rng(3)
    n, p = 1 << 16, 3                                   # raw IQ samples, /8
    iq = rng.integers(0, 256, 2 * n, dtype=np.uint8)
    lead = 64 << p
    lo, hi = rank * n // world, (rank + 1) * n // world
    lo_in = lo - lead if rank else lo
    part = ol.downconvert(iq[2 * lo_in:2 * hi], p, 0)[(2 * 64 if rank else 0):]
    parts = [None] * world
    dist.all_gather_object(parts, part.tobytes())
    if rank == 0:
        out.put((gathered, b"".join(parts) == ol.downconvert(iq, p, 0).tobytes()))
    dist.destroy_process_group()


def test_two_ranks_decode_their_streams_and_time_shard_a_buffer():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import iqsynth as g
    import oracle_lib as ol
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_decode_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, same = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "time-sharded decimation differs from the whole-buffer result"
    got = {}
    for d in gathered:
        got.update(d)
    assert sorted(got) == [0, 1, 2, 3, 4]
    for s_ in range(5):                                  # what one process decodes for every stream
        o = ol.Oracle(types=0x07, thresh=0)
        o.process(g.fixture_single_tfa1(seed=1 + s_))
        assert got[s_] == [r["exec"] for r in o.records()] and len(got[s_]) == 1
        o.close()
