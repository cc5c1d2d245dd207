#!/usr/bin/env python
"""Does cutting a device-resident trajectory scan into pose chunks pay on ONE GPU?  The ordered compaction (HBM-bound, 9 % of
the C2 step) of chunk c then runs on the auxiliary stream while chunk c+1 is traversed (k_trace leaves DRAM 96 % idle).

    python tools/chunk_overlap.py [--workload c2] [--plans 1,2,3,4,6] [--taper 1,3]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lrc_b200 as lrc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--plans", default="1,2,3,4,6")
    ap.add_argument("--taper", default="1,3")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    w, mesh, poses, intr = bench.make_workload(lrc, args.workload, 1)
    dev = torch.device("cuda", 0)
    ctx = lrc.RaycastEngineGPU(device=0).ctx
    v, f, lab = lrc.mesh_arrays(mesh)
    ctx.set_mesh_arrays(v, f, lab)
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2) if w["noise"] else None
    n_frame, P = lrc.rays_per_frame(intr), len(poses)
    poses_d = torch.from_numpy(np.ascontiguousarray(poses.reshape(-1, 16))).to(dev)
    bufs, _ = ctx._alloc_out(P * n_frame, P)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for chunks in [int(x) for x in args.plans.split(",")]:
        for taper in [int(x) for x in args.taper.split(",")]:
            if chunks == 1 and taper != 1:
                continue
            ctx.set_option("scan_chunks", chunks)
            ctx.set_option("scan_taper", taper)
            ts = []
            for r in range(args.reps + 3):
                flush.fill_(r & 255)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ctx.scan_enqueue(poses_d, intr, noise, bufs)
                b.record()
                torch.cuda.synchronize()
                if r >= 3:
                    ts.append(a.elapsed_time(b))
            m = int(bufs["off"][-1].item())
            sig = (m, int(bufs["prim"][:m].to(torch.int64).sum().item()), float(bufs["xyz"][:m].double().sum().item()),
                   int(bufs["off"].sum().item()))
            ref = ref or sig
            print(json.dumps({"workload": args.workload, "chunks": chunks, "taper": taper, "step_ms": round(float(np.mean(ts)), 4),
                              "step_ms_min": round(float(np.min(ts)), 4), "Mrays_s": round(P * n_frame / np.mean(ts) / 1e3, 1),
                              "same_output": sig == ref}), flush=True)
    ctx.set_option("scan_chunks", 1)


if __name__ == "__main__":
    main()
