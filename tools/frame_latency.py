#!/usr/bin/env python
"""Latency of the reference's own call pattern: one `engine.lidar_intersect_mesh(lidar, mesh)` per waypoint
(s3dis_simulator.py:254-263), numpy in / numpy out, BVH cached across calls.  Two cache modes: the mesh pinned with
`engine.set_mesh(mesh)` (identity check only) and the default (a CRC over the whole mesh on every call).

    python tools/frame_latency.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrc_b200 as lrc  # noqa: E402


def per_frame_latency(eng, mesh, intr, poses, pinned: bool = True, warm: int = 5) -> dict:
    """Median / p90 wall time of `lidar_intersect_mesh` over `poses` (after `warm` untimed calls), as a user of the
    two-line swap in INTEGRATION.md sees it: create_lidar + engine call + fresh numpy arrays back."""
    if pinned:
        eng.set_mesh(mesh)
    else:
        eng.invalidate_mesh()
    for p in poses[:warm]:
        eng.lidar_intersect_mesh(lrc.create_lidar(intr, p), mesh)
    ts, pts = [], None
    for p in poses[warm:]:
        t0 = time.perf_counter()
        lidar = lrc.create_lidar(intr, p)
        pts, inc = eng.lidar_intersect_mesh(lidar, mesh)
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    n = lrc.rays_per_frame(intr)
    return {"ms_median": float(np.median(ts)), "ms_p90": float(np.percentile(ts, 90)), "frames": int(len(ts)),
            "frames_per_s": float(1e3 / np.median(ts)), "Mrays_per_s": float(n / np.median(ts) / 1e3),
            "rays_per_frame": int(n), "points_last_frame": 0 if pts is None else int(len(pts)), "mesh_cache": "pinned" if pinned else "content-checked"}


def main():
    eng = lrc.RaycastEngineGPU()
    mesh = lrc.synthetic.office()
    poses = lrc.poses_from_waypoints(lrc.synthetic.office_waypoints(60))
    sensors = (("8-line (16 000 rays)", lrc.Indoor8LineLidarIntrinsics.create_standard_8line()),
               ("32-line (128 000 rays)", lrc.Indoor8LineLidarIntrinsics.create_dense_32line()),
               ("BLK2GO dual-axis (64 000 rays)", lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()))
    for pinned in (True, False):
        for name, intr in sensors:
            r = per_frame_latency(eng, mesh, intr, poses, pinned=pinned)
            print(f"[{r['mesh_cache']}] {name}: median {r['ms_median']:.3f} ms per frame ({r['frames_per_s']:.0f} frames/s, "
                  f"{r['Mrays_per_s']:.0f} Mrays/s), p90 {r['ms_p90']:.3f} ms, {r['points_last_frame']} points in the last frame", flush=True)


if __name__ == "__main__":
    main()
