#!/usr/bin/env python
"""Aggregate device-to-host rate of N ranks copying at once -- what bounds `e2e` at N > 1 (every rank returns its own
307 MB cloud per step).  Plain cudaMemcpyAsync from device memory into page-locked host memory, no engine involved.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe_multi.py

Per phase: (a) every rank alone in turn (the others idle), (b) all ranks at once.  Prints one JSON line on rank 0.
"""
import json
import os

import torch
import torch.distributed as dist


def timed_copy(h, d, reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 307_198_736                     # bytes a C2 step returns per rank: 12.8 M points x 24 B
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    timed_copy(h, d, 2)
    alone = torch.zeros(world, dtype=torch.float64, device="cuda")
    for r in range(world):
        dist.barrier()
        if r == rank:
            alone[r] = n / timed_copy(h, d, 5) / 1e6
        dist.barrier()
    dist.all_reduce(alone)
    dist.barrier()
    together = torch.zeros(world, dtype=torch.float64, device="cuda")
    together[rank] = n / timed_copy(h, d, 8) / 1e6
    dist.all_reduce(together)
    if rank == 0:
        print(json.dumps({"ranks": world, "bytes_per_copy": n, "alone_gbs_per_rank": [round(x, 1) for x in alone.tolist()],
                          "together_gbs_per_rank": [round(x, 1) for x in together.tolist()],
                          "together_aggregate_gbs": round(float(together.sum()), 1),
                          "cpus_visible": len(os.sched_getaffinity(0))}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
