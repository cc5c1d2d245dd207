#!/usr/bin/env python
"""Latency of the reference's own call pattern: one `engine.lidar_intersect_mesh(lidar, mesh)` per waypoint
(s3dis_simulator.py:254-263), numpy in / numpy out, BVH cached across calls.

    python tools/frame_latency.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrc_b200 as lrc  # noqa: E402


def main():
    eng = lrc.RaycastEngineGPU()
    mesh = lrc.synthetic.office()
    wps = lrc.synthetic.office_waypoints(60)
    for name, intr in (("8-line (16 000 rays)", lrc.Indoor8LineLidarIntrinsics.create_standard_8line()),
                       ("32-line (128 000 rays)", lrc.Indoor8LineLidarIntrinsics.create_dense_32line()),
                       ("BLK2GO dual-axis (64 000 rays)", lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis())):
        eng.lidar_intersect_mesh(lrc.create_lidar(intr, wps[0].to_pose_matrix()), mesh)      # builds the LBVH
        ts = []
        for w in wps[10:]:
            lidar = lrc.create_lidar(intr, w.to_pose_matrix())
            t0 = time.perf_counter()
            pts, inc = eng.lidar_intersect_mesh(lidar, mesh)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print(f"{name}: median {np.median(ts):.3f} ms per frame ({1e3 / np.median(ts):.0f} frames/s), p90 {np.percentile(ts, 90):.3f} ms, "
              f"{len(pts)} points in the last frame", flush=True)


if __name__ == "__main__":
    main()
