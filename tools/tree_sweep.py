#!/usr/bin/env python
"""Round 2: tree builder x traversal variant x leaf size on one workload.

    python tools/tree_sweep.py [--workload c2] [--reps 8] [--quality 0,1] [--variants 1,65] [--leaf 2,4] [--radius 16]

Per combination: LBVH / PLOC build time (CUDA events, second build), node and triangle records fetched per ray
(counting instantiation), tree height, SAH cost, device time of k_trace and of the compaction (CUDA events inside the
library, L2 flushed before every repetition) and a signature of the output that must not change (the closest hit is
independent of the tree).
"""
import argparse
import itertools
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
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--quality", default="0,1")
    ap.add_argument("--variants", default="1,65")
    ap.add_argument("--leaf", default="2")
    ap.add_argument("--radius", default="8")
    ap.add_argument("--formats", default="0", help="node formats: 0 = 64 B centre/half, 1 = 32 B 16-bit, 2 = 64 B paired (packed FMA)")
    ap.add_argument("--persistent", default="0", help="0 / 1: persistent warps over 32-ray tiles")
    ap.add_argument("--rpt", default="1", help="rays per thread (1, 2, 4); > 1 needs format 2")
    ap.add_argument("--block", default="128", help="threads per traversal block (32 / 64 / 128)")
    ap.add_argument("--wp", default="0", help="0 / 1: warp-packet traversal (format 2)")
    ap.add_argument("--tune", default="0", help="format-2 kernel tuning bits: 1 prefetch, 2 no block barrier, 4 streaming scratch stores")
    ap.add_argument("--poses", type=int, default=None)
    ap.add_argument("--tris", type=int, default=None)
    args = ap.parse_args()
    w, mesh, poses, intr = bench.make_workload(lrc, args.workload, 1, args.tris, args.poses)
    dev = torch.device("cuda", 0)
    ctx = lrc.RaycastEngineGPU(device=0).ctx
    v, f, lab = lrc.mesh_arrays(mesh)
    v_d, f_d = torch.from_numpy(v).to(dev), torch.from_numpy(f).to(dev)
    l_d = torch.from_numpy(lab.view(np.int32)).to(dev)
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2) if w["noise"] else None
    n_frame = lrc.rays_per_frame(intr)
    P = len(poses)
    poses_d = torch.from_numpy(np.ascontiguousarray(poses.reshape(-1, 16))).to(dev)
    bufs, _ = ctx._alloc_out(P * n_frame, P)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    ints = lambda s: [int(x) for x in s.split(",")]
    for q, leaf, rad, fmt in itertools.product(ints(args.quality), ints(args.leaf), ints(args.radius), ints(args.formats)):
        if q == 0 and rad != ints(args.radius)[0]:
            continue
        ctx.set_option("node_format", fmt)
        ctx.set_option("build_quality", q)
        ctx.set_option("leaf_size", leaf)
        ctx.set_option("ploc_radius", rad)
        ctx.set_mesh_arrays(v_d, f_d, l_d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.set_mesh_arrays(v_d, f_d, l_d)
        e1.record()
        torch.cuda.synchronize()
        build_ms = e0.elapsed_time(e1)
        info = ctx.bvh_info()
        for var, pers, rpt, tune, wp in itertools.product(ints(args.variants), ints(args.persistent), ints(args.rpt), ints(args.tune), ints(args.wp)):
          for blk in ints(args.block):
                if rpt > 1 and (fmt != 2 or pers or var != ints(args.variants)[-1]):
                    continue
                if tune and (fmt != 2 or pers or rpt > 1 or var != 65):
                    continue
                if wp and (fmt != 2 or pers or rpt > 1 or tune):
                    continue
                ctx.set_option("warp_packet", wp)
                ctx.set_option("block", blk)
                ctx.set_option("tune", tune)
                ctx.set_option("variant", var)
                ctx.set_option("persistent", pers)
                ctx.set_option("rays_per_thread", rpt)
                ctx.set_counting(True)
                ctx.counters(reset=True)
                ctx.scan_enqueue(poses_d, intr, noise, bufs)
                cnt = ctx.counters(reset=True)
                ctx.set_counting(False)
                ctx.set_option("kernel_timing", 1)
                tr, cp = [], []
                for r in range(args.reps + 2):
                    flush.fill_(r & 255)
                    ctx.scan_enqueue(poses_d, intr, noise, bufs)
                    torch.cuda.synchronize()
                    kt = ctx.kernel_times()
                    if r >= 2:
                        tr.append(kt["trace_ms"])
                        cp.append(kt["compact_ms"])
                ctx.set_option("kernel_timing", 0)
                m = int(bufs["off"][-1].item())
                sig = (m, int(bufs["prim"][:m].to(torch.int64).sum().item()), float(bufs["xyz"][:m].double().sum().item()))
                if ref is None:
                    ref = sig
                rays = max(1, cnt["rays"])
                print(json.dumps({"workload": args.workload, "quality": q, "leaf_size": leaf, "ploc_radius": rad if q else None, "node_format": fmt, "variant": var, "persistent": pers, "rays_per_thread": rpt, "tune": tune, "warp_packet": wp, "block": blk,
                                  "build_ms": round(build_ms, 3), "height": info["max_depth"], "sah": round(info["sah_cost"], 2),
                                  "nodes_per_ray": round(cnt["nodes_visited"] / rays, 2), "tris_per_ray": round(cnt["tris_tested"] / rays, 2),
                                  "trace_ms": round(float(np.mean(tr)), 4), "trace_ms_min": round(float(np.min(tr)), 4),
                                  "compact_ms": round(float(np.mean(cp)), 4), "Mrays_s_trace": round(P * n_frame / np.mean(tr) / 1e3, 1),
                                  "same_output": sig == ref}), flush=True)


if __name__ == "__main__":
    main()
