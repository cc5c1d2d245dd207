#!/usr/bin/env python
"""Time k_trace / compaction for every traversal variant x block size on one workload (CUDA events inside the
library, L2 flushed before every run) and check that all variants produce identical outputs.

    python tools/variant_sweep.py [--workload c2] [--reps 10]
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
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="0,1,2,3")
    ap.add_argument("--blocks", default="32,64,128")
    ap.add_argument("--formats", default="0", help="comma list of node formats (0 = 64 B float, 1 = 32 B 16-bit); format 1 ignores --variants")
    ap.add_argument("--leaf", default="1", help="comma list of leaf sizes (triangles per leaf, rebuilds the tree)")
    ap.add_argument("--l2", default="0", help="comma list of l2_persist percentages")
    ap.add_argument("--stack", default="12", help="comma list of stack_levels for the shared-memory-stack variants (bit 4)")
    ap.add_argument("--top", default="6", help="comma list of top_levels for the shared-memory variants (bit 3)")
    args = ap.parse_args()
    w, mesh, poses, intr = bench.make_workload(lrc, args.workload, 1)
    dev = torch.device("cuda", 0)
    eng = lrc.RaycastEngineGPU(device=0)
    ctx = eng.ctx
    v, f, lab = lrc.mesh_arrays(mesh)
    ctx.set_mesh_arrays(v, f, lab)
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2) if w["noise"] else None
    n_frame = lrc.rays_per_frame(intr)
    P = len(poses)
    poses_d = torch.from_numpy(np.ascontiguousarray(poses.reshape(-1, 16))).to(dev)
    bufs, _ = ctx._alloc_out(P * n_frame, P)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    ctx.set_option("kernel_timing", 1)
    import itertools
    combos = []
    for fmt in [int(x) for x in args.formats.split(",")]:
        if fmt == 1:
            combos.append((37, 0, 0, 0))
    for var, l2 in itertools.product([] if args.formats == "1" else [int(x) for x in args.variants.split(",")], [int(x) for x in args.l2.split(",")]):
        for top in ([int(x) for x in args.top.split(",")] if var & 8 else [0]):
            for stk in ([int(x) for x in args.stack.split(",")] if var & 16 else [0]):
                combos.append((var, l2, top, stk))
    cur_fmt = (0, 1)
    combos = [c + (lf,) for lf in [int(x) for x in args.leaf.split(",")] for c in combos]
    for var, l2, top, stk, leaf in combos:
        fmt = 1 if var == 37 else 0
        if (fmt, leaf) != cur_fmt:
            ctx.set_option("node_format", fmt)
            ctx.set_option("leaf_size", leaf)
            ctx.set_mesh_arrays(v, f, lab)
            cur_fmt = (fmt, leaf)
        if var == 37:
            var_opt = 5
        else:
            var_opt = var
        if top:
            ctx.set_option("top_levels", top)
        if stk:
            ctx.set_option("stack_levels", stk)
        for blk in [int(x) for x in args.blocks.split(",")]:
            ctx.set_option("variant", var_opt)
            ctx.set_option("block", blk)
            ctx.set_option("l2_persist", l2)
            tr, cp = [], []
            for r in range(args.reps + 2):
                ctx.set_option("l2_reset", 1)
                flush.fill_(r & 255)
                ctx.scan_enqueue(poses_d, intr, noise, bufs)
                torch.cuda.synchronize()
                kt = ctx.kernel_times()
                if r >= 2:
                    tr.append(kt["trace_ms"])
                    cp.append(kt["compact_ms"])
            m = int(bufs["off"][-1].item())
            sig = (m, int(bufs["prim"][:m].to(torch.int64).sum().item()), float(bufs["xyz"][:m].double().sum().item()))
            if ref is None:
                ref = sig
            print(json.dumps({"variant": var, "block": blk, "l2_persist": l2, "top_levels": top, "stack_levels": stk, "leaf_size": leaf, "trace_ms": round(float(np.mean(tr)), 4),
                              "trace_ms_min": round(float(np.min(tr)), 4), "compact_ms": round(float(np.mean(cp)), 4),
                              "Mrays_s_trace": round(P * n_frame / np.mean(tr) / 1e3, 1), "same_output": sig == ref}), flush=True)


if __name__ == "__main__":
    main()
