"""Bare host->device copy ceiling of the node at N ranks: what bench.py's `e2e` can reach at best.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_probe.py

Every rank pins one buffer (cudaHostAlloc through torch), copies it to its GPU `reps` times with cudaMemcpyAsync on
one stream and times the copies with CUDA events: first all ranks at once (what the bench does), then one rank after
the other.  Rank 0 prints one JSON line.  No kernels of the repo are involved: this is the platform."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    host = torch.empty(mib << 20, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)
    devb = torch.empty(mib << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def copy_gbs():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        devb.copy_(host, non_blocking=True)          # warm-up
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            devb.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        return reps * (mib << 20) / (e0.elapsed_time(e1) / 1e3) / 1e9

    barrier()
    together = copy_gbs()
    barrier()
    alone = 0.0
    for r in range(world):
        if r == rank:
            alone = copy_gbs()
        barrier()
    vals = torch.tensor([together, alone], dtype=torch.float64, device=dev)
    if world > 1:
        allv = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        tg = [round(float(v[0]), 2) for v in allv]
        al = [round(float(v[1]), 2) for v in allv]
        print(json.dumps({"probe": "pinned host -> device cudaMemcpyAsync", "n_gpus": world, "mib_per_copy": mib, "reps": reps,
                          "gbs_per_rank_all_at_once": tg, "gbs_per_rank_alone": al,
                          "sum_all_at_once": round(sum(tg), 2), "host_cores": len(os.sched_getaffinity(0))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
