"""
Golden fixtures for the post-cast steps, produced by the LIVE REFERENCE at /root/reference (run once in the build
container; the reference tree does not exist on the GPU box):

    python tests/golden/make_golden_post.py

  post_labeled.ply   written by the reference's own S3DISSimScene._save_labeled_ply
                     (containers/s3dis_sim_scene.py:614-641) for 257 seeded points
  post_stats.npz     the inputs above plus: ScanQuality objects built by the reference's dataclass from the
                     expressions of s3dis_simulator.py:276-284 evaluated here on three seeded frames (one empty),
                     and the SimulationStats the reference's S3DISSimScene.compute_statistics derives from them
The reference's ``containers`` package imports with a stub ``open3d`` (only s3dis_scene.py needs the name).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
_o3d = types.ModuleType("open3d")
_o3d.geometry = types.SimpleNamespace(TriangleMesh=object, PointCloud=object, AxisAlignedBoundingBox=object)
sys.modules.setdefault("open3d", _o3d)
sys.path.insert(0, REF)
from containers.s3dis_sim_frame import S3DISSimFrame, ScanQuality  # noqa: E402
from containers.s3dis_sim_scene import S3DISSimScene  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    # ---- PLY ----
    n = 257                                                     # not a multiple of the 256-point packing tile
    pts = (rng.standard_normal((n, 3)) * 7.0).astype(np.float32)
    pts[0] = (0.0, -0.0, 1e-30)
    pts[1] = (np.float32(3.4e38), np.float32(-1.17549435e-38), np.float32(123456.789))
    colors = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    sem = rng.integers(0, 13, n).astype(np.uint16)
    ins = rng.integers(0, 65536, n).astype(np.uint16)
    scene = S3DISSimScene("golden")
    path = os.path.join(HERE, "post_labeled.ply")
    from pathlib import Path
    scene._save_labeled_ply(Path(path), pts, colors, sem, ins)
    print("post_labeled.ply", os.path.getsize(path), "bytes")

    # ---- ScanQuality / SimulationStats ----
    total_points_per_scan, room_volume = 16000, 10.0 * 8.0 * 3.0
    out = {"ply/points": pts, "ply/colors": colors, "ply/sem": sem, "ply/ins": ins,
           "total_points_per_scan": np.int64(total_points_per_scan), "room_volume": np.float64(room_volume)}
    scene = S3DISSimScene("golden")
    sizes = [5000, 0, 12345]
    for i, m in enumerate(sizes):
        points = (rng.standard_normal((m, 3)) * 4.0 + np.array([3.0, 2.5, 1.0])).astype(np.float32)
        incident_angles = rng.uniform(0.0, 90.0, m)
        # the expressions of s3dis_simulator.py:276-284, fed to the reference's own dataclass
        q = ScanQuality(
            coverage_ratio=len(points) / total_points_per_scan,
            num_points=len(points),
            incident_angle_mean=np.mean(incident_angles) if len(incident_angles) > 0 else 0,
            incident_angle_std=np.std(incident_angles) if len(incident_angles) > 0 else 0,
            scan_density=len(points) / room_volume,
            range_mean=np.mean(np.linalg.norm(points, axis=1)) if len(points) > 0 else 0,
            range_std=np.std(np.linalg.norm(points, axis=1)) if len(points) > 0 else 0)
        scene.append_frame(S3DISSimFrame(i, points, incident_angles, q))
        out[f"frame{i}/points"], out[f"frame{i}/incident"] = points, incident_angles
        d = q.to_dict()
        out[f"frame{i}/quality"] = np.array([d[k] for k in ("coverage_ratio", "num_points", "incident_angle_mean", "incident_angle_std",
                                                            "scan_density", "range_mean", "range_std")], dtype=np.float64)
    scene.compute_statistics(simulation_time=2.5)
    s = scene.statistics.to_dict()
    out["stats"] = np.array([s[k] for k in ("total_frames", "total_points", "average_coverage", "average_scan_density",
                                            "average_incident_angle", "average_range", "simulation_time", "frames_per_second")], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "post_stats.npz"), **out)
    print("post_stats.npz", os.path.getsize(os.path.join(HERE, "post_stats.npz")), "bytes")


if __name__ == "__main__":
    main()
