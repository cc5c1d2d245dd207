#!/usr/bin/env python
"""Block size of the traversal kernel for SMALL launches (one frame per call): a 16 000-ray frame is 125 blocks of 128
threads on 148 SMs -- one warp per scheduler, nothing to hide latency with.  Device time of one frame (CUDA events around
the library's launch sequence, device-resident pose) and wall time of the per-waypoint call, per block size."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrc_b200 as lrc  # noqa: E402


def main():
    eng = lrc.RaycastEngineGPU()
    ctx = eng.ctx
    mesh = lrc.synthetic.office()
    eng.set_mesh(mesh)
    poses = lrc.poses_from_waypoints(lrc.synthetic.office_waypoints(60))
    sensors = (("8-line", lrc.Indoor8LineLidarIntrinsics.create_standard_8line()),
               ("blk2go", lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()),
               ("32-line", lrc.Indoor8LineLidarIntrinsics.create_dense_32line()))
    for name, intr in sensors:
        n = lrc.rays_per_frame(intr)
        bufs, _ = ctx._alloc_out(n, 1)
        for blk in (128, 64, 32, 0):
            ctx.set_option("block", blk)
            dev_ms, wall_ms = [], []
            for k, p in enumerate(poses):
                pd = torch.from_numpy(np.ascontiguousarray(p.reshape(1, 16))).cuda()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ctx.scan_enqueue(pd, intr, None, bufs); b.record(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                eng.lidar_intersect_mesh(lrc.create_lidar(intr, p), mesh)
                t1 = time.perf_counter()
                if k >= 10:
                    dev_ms.append(a.elapsed_time(b)); wall_ms.append((t1 - t0) * 1e3)
            print(f"{name} ({n} rays) block {blk if blk else 'auto'}: device sequence {np.median(dev_ms):.4f} ms, per-waypoint call {np.median(wall_ms):.4f} ms", flush=True)
    ctx.set_option("block", 0)


if __name__ == "__main__":
    main()
