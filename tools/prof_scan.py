#!/usr/bin/env python
"""A few device-resident trajectory scans with chosen engine options -- the command profiled under ncu.

    python tools/prof_scan.py --workload c2 [--format 2 --variant 65 --tune 2 --rpt 1 --quality 0 --wp 0] --reps 3

Options that are not given keep the library's defaults (the build bench.py runs).  Prints the build tag bench.py
matches a capture against (profiles/traffic.json).
"""
import argparse
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
    ap.add_argument("--format", type=int, default=None)
    ap.add_argument("--variant", type=int, default=None)
    ap.add_argument("--tune", type=int, default=None)
    ap.add_argument("--wp", type=int, default=None)
    ap.add_argument("--rpt", type=int, default=None)
    ap.add_argument("--quality", type=int, default=None)
    ap.add_argument("--leaf", type=int, default=None)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--poses", type=int, default=None)
    ap.add_argument("--tris", type=int, default=None)
    args = ap.parse_args()
    w, mesh, poses, intr = bench.make_workload(lrc, args.workload, 1, args.tris, args.poses)
    dev = torch.device("cuda", 0)
    ctx = lrc.RaycastEngineGPU(device=0).ctx
    for k, v in (("node_format", args.format), ("variant", args.variant), ("tune", args.tune), ("warp_packet", args.wp),
                 ("rays_per_thread", args.rpt), ("build_quality", args.quality), ("leaf_size", args.leaf)):
        if v is not None:
            ctx.set_option(k, v)
    v, f, lab = lrc.mesh_arrays(mesh)
    ctx.set_mesh_arrays(v, f, lab)
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2) if w["noise"] else None
    n_frame = lrc.rays_per_frame(intr)
    P = len(poses)
    poses_d = torch.from_numpy(np.ascontiguousarray(poses.reshape(-1, 16))).to(dev)
    bufs, _ = ctx._alloc_out(P * n_frame, P)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for r in range(args.reps):
        flush.fill_(r & 255)
        ctx.scan_enqueue(poses_d, intr, noise, bufs)
        torch.cuda.synchronize()
    print(json_line := {"workload": args.workload, "tris": int(len(f)), "rays_per_launch": int(P * n_frame), "build_tag": bench.build_tag(ctx),
                        "points": int(bufs["off"][-1].item())})


if __name__ == "__main__":
    main()
