#!/usr/bin/env python
"""All-to-all NVLink probe (run under torchrun, one rank per GPU): how fast can every rank deliver the same `--mb` MB
block to all its peers at once (a) with copy engines -- one cudaMemcpyAsync per peer on its own stream -- and (b) with
the engine's SM-driven exchange kernel (k_push, measured through a real gathered scan elsewhere)?  Prints, per rank-0,
the aggregate egress bandwidth of (a) while all ranks send simultaneously."""
import argparse
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=51.2)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lrc_b200 as lrc
    from lrc_b200.distributed import PeerGather
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = lrc.get_context(local)
    n = int(args.mb * 1e6) // 16                      # "points" of 16 bytes
    pg = PeerGather(ctx, cap_per_rank=n, frames_per_rank=1)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    src = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(world)]
    nbytes = n * 12
    for mode in ("ce_all_peers", "ce_one_peer"):
        dist.barrier()
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(args.reps + 2):
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            peers = [r for r in range(world) if r != rank]
            if mode == "ce_one_peer":
                peers = [(rank + 1) % world]
            for r in peers:
                streams[r].wait_event(e0)
                rc = rt.cudaMemcpyAsync(C.c_void_p(pg.ptrs[r] + rank * n * 12), C.c_void_p(src.data_ptr()), nbytes, 3,
                                        C.c_void_p(streams[r].cuda_stream))
                assert rc == 0, rc
                ev = torch.cuda.Event()
                ev.record(streams[r])
                torch.cuda.current_stream().wait_event(ev)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 2:
                best = min(best, e0.elapsed_time(e1))
        t = torch.tensor([best], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            gb = nbytes * len(peers) / 1e9
            print(f"{mode}: {gb * 1e3:.0f} MB egress per rank in {t.item():.3f} ms (max over ranks) -> {gb / (t.item() * 1e-3):.0f} GB/s per rank", flush=True)
    pg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
