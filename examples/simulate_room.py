#!/usr/bin/env python
"""End-to-end room simulation on the GPU, the way the reference's `python s3dis_simulator.py` does it for one scene
(s3dis_simulator.py:407-444): plan a trajectory -> ray-cast every waypoint -> scan statistics -> labelled PLY.

    python examples/simulate_room.py [--tris 200000] [--waypoints 20] [--sensor blk2go|32line|8line] [--out /tmp/room]

Needs a B200 and the built library (`python __graft_entry__.py`).  Uses a synthetic furnished room (S3DIS / NKSR meshes
are not redistributable); any object with `.vertices` / `.triangles` works as the mesh.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrc_b200 as lrc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tris", type=int, default=200_000)
    ap.add_argument("--waypoints", type=int, default=20)
    ap.add_argument("--sensor", default="blk2go", choices=["blk2go", "32line", "8line"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="/tmp/lrc_room")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    mesh = lrc.synthetic.box_room(target_tris=args.tris, seed=3)
    bounds = lrc.room_bounds_of(mesh)
    lidar_config = {"blk2go": lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis,
                    "32line": lrc.Indoor8LineLidarIntrinsics.create_dense_32line,
                    "8line": lrc.Indoor8LineLidarIntrinsics.create_standard_8line}[args.sensor]()
    engine = lrc.RaycastEngineGPU()

    np.random.seed(args.seed)                         # the planner draws its candidates from the global numpy stream
    t0 = time.perf_counter()
    waypoints, info = lrc.AutoTrajectoryGenerator().generate_optimal_trajectory(mesh, bounds, num_waypoints=args.waypoints)
    t_plan = time.perf_counter() - t0
    print(f"trajectory: {len(waypoints)} waypoints, {info['best_trajectory']['length']:.2f} m, "
          f"{info['best_trajectory']['collision_count']} collisions, {info['total_candidates']} candidates, "
          f"{info['room_analysis']['free_space_points']} free samples  [{t_plan * 1e3:.1f} ms]")

    engine.set_mesh(mesh)                             # LBVH build (cached for the run below)
    run = lrc.run_simulation(engine, waypoints, lidar_config, mesh,
                             noise=lrc.NoiseConfig.from_intrinsics(lidar_config, seed=args.seed), bounds=bounds)
    s = run.statistics
    print(f"simulation: {s.total_frames} frames, {s.total_points:,} points, coverage {s.average_coverage:.3f}, "
          f"incident {s.average_incident_angle:.1f} deg, range {s.average_range:.2f} m, "
          f"{s.simulation_time * 1e3:.1f} ms -> {s.frames_per_second:.0f} frames/s")
    for f in run.frames[:3]:
        q = f.scan_quality
        print(f"  frame {f.frame_index}: {q.num_points} points, coverage {q.coverage_ratio:.3f}, incident {q.incident_angle_mean:.1f} +- {q.incident_angle_std:.1f}")
    path = os.path.join(args.out, "combined_pointcloud_with_label.ply")
    n = lrc.write_labeled_ply(engine.ctx, path, run.scan)
    back = lrc.read_labeled_ply(path)
    sem, counts = np.unique(back["semantic_labels"], return_counts=True)
    print(f"wrote {path}: {n / 1e6:.1f} MB, semantic classes {dict(zip(sem.tolist(), counts.tolist()))}")


if __name__ == "__main__":
    main()
