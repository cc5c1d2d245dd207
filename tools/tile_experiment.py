#!/usr/bin/env python
"""Does the shape of a warp's ray tile matter?  Cast the C2 rays (explicit ray buffer, dense outputs) in different
orders: 32 consecutive azimuths of one scan line per warp (the engine's order), 16 x 2 lines, 8 x 4 lines, and a random
permutation (incoherent upper bound).  CUDA events around lrc_cast_rays, L2 flushed before every run.

    python tools/tile_experiment.py [--poses 20]
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
    ap.add_argument("--poses", type=int, default=20)
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    w, mesh, poses, intr = bench.make_workload(lrc, args.workload, 1, None, args.poses)
    dev = torch.device("cuda", 0)
    ctx = lrc.RaycastEngineGPU(device=0).ctx
    v, f, lab = lrc.mesh_arrays(mesh)
    ctx.set_mesh_arrays(v, f, lab)
    rays, _ = ctx.gen_rays(poses, intr)
    H, W = len(intr.vertical_degrees), intr.horizontal_res
    P = len(poses)
    n = rays.shape[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    idx = torch.arange(n, device=dev).reshape(P, H, W)
    orders = {"32x1 (engine order)": idx.reshape(-1)}
    for a, l in ((16, 2), (8, 4), (4, 8)):
        if H % l == 0 and W % a == 0:
            t = idx.reshape(P, H // l, l, W // a, a).permute(0, 1, 3, 2, 4)      # [P, H/l, W/a, l, a]
            orders[f"{a}x{l}"] = t.reshape(-1)
    orders["random"] = torch.randperm(n, device=dev)
    base = None
    for name, perm in orders.items():
        r = rays[perm].contiguous()
        ts = []
        for k in range(args.reps + 2):
            flush.fill_(k & 255)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            t, pid = ctx.cast_rays(r)
            b.record()
            torch.cuda.synchronize()
            if k >= 2:
                ts.append(a.elapsed_time(b))
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(n, device=dev)
        sig = int(pid[inv].to(torch.int64).sum().item())
        base = sig if base is None else base
        ctx.set_counting(True)
        ctx.counters(reset=True)
        ctx.cast_rays(r)
        c = ctx.counters(reset=True)
        ctx.set_counting(False)
        print(json.dumps({"order": name, "ms": round(float(np.mean(ts)), 4), "Mrays_s": round(n / np.mean(ts) / 1e3, 1),
                          "nodes_per_ray": round(c["nodes_visited"] / c["rays"], 2), "same_hits": sig == base}), flush=True)


if __name__ == "__main__":
    main()
