"""
Generates the committed golden fixtures in this directory FROM THE LIVE REFERENCE at /root/reference.
Run once in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

What is captured
  rays_*.npz      ray tables of the reference's own ``lidar`` package (numpy-only, importable):
                  IndoorLidar.get_rays for the 8/32/64-line presets and the uniform-fov branch,
                  DualAxisLidar.get_rays with angle noise and dropout zeroed, each at the identity pose and
                  at a translated + yawed pose.  Stored: every STRIDE-th ray, plus sha256 / float64 sums of
                  the full table.
  poses.npz       Waypoint.to_pose_matrix of the reference's ``trajectory`` package (imported with a stub
                  ``open3d`` module: only a type annotation needs it).
  frame_*.npz     the reference's UNMODIFIED RaycastEngineCPU.lidar_intersect_mesh /
                  rays_intersect_mesh (raycast_engine/raycast_engine_cpu.py) executed with a stub ``open3d``
                  whose RaycastingScene.cast_rays is backed by oracle/liblrc_oracle.so.  This pins the
                  reference's numpy epilogue (point reconstruction, range filter, incident angles, ordered
                  compaction) -- everything except the third-party intersector itself, which is absent.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
STRIDE = 37

sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

# ---- stub open3d: only what the reference's engine touches ----------------------------------------
o3d = types.ModuleType("open3d")


class _Tensor:
    def __init__(self, a):
        self._a = np.asarray(a)

    def numpy(self):
        return self._a


class _LegacyMesh:
    def __init__(self, vertices, triangles):
        self.vertices, self.triangles = vertices, triangles


class _TMesh:
    @staticmethod
    def from_legacy(mesh):
        return mesh


class _Scene:
    def add_triangles(self, mesh):
        self._scene = orc.OracleScene(mesh)

    def cast_rays(self, rays):
        t, pid = self._scene.cast_rays(rays.numpy())
        self.last_prim = pid
        _Scene.last = self
        return {"t_hit": _Tensor(t), "primitive_normals": _Tensor(np.zeros((len(t), 3), np.float32)),
                "primitive_ids": _Tensor(pid)}


o3d.geometry = types.SimpleNamespace(TriangleMesh=_LegacyMesh, PointCloud=object, AxisAlignedBoundingBox=object)
o3d.t = types.SimpleNamespace(geometry=types.SimpleNamespace(RaycastingScene=_Scene, TriangleMesh=_TMesh))
o3d.core = types.SimpleNamespace(Tensor=_Tensor)
o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: a, Vector3iVector=lambda a: a)
o3d.io = types.SimpleNamespace()
o3d.visualization = types.SimpleNamespace()
sys.modules["open3d"] = o3d

# reference imports: the engine directory first so that `from raycast_engine import RaycastEngineBase`
# (the except-ImportError branch of raycast_engine_cpu.py:12-14) finds raycast_engine/raycast_engine.py
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "raycast_engine"))
from lidar import (DualAxisLidar, DualAxisLidarIntrinsics, Indoor8LineLidarIntrinsics, IndoorLidar,  # noqa: E402
                   create_lidar)

spec = importlib.util.spec_from_file_location("ref_raycast_engine_cpu", os.path.join(REF, "raycast_engine", "raycast_engine_cpu.py"))
ref_cpu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_cpu)

spec = importlib.util.spec_from_file_location("ref_trajectory_generator", os.path.join(REF, "trajectory", "trajectory_generator.py"))
ref_traj = importlib.util.module_from_spec(spec)
sys.modules["ref_trajectory_generator"] = ref_traj
spec.loader.exec_module(ref_traj)


def pose_of(x, y, z, yaw):
    return ref_traj.Waypoint(x, y, z, yaw).to_pose_matrix()


POSES = {"identity": np.eye(4), "posed": pose_of(3.137, 2.718, 1.0, 0.3)}


def table_record(rays):
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    return {
        "n": np.int64(len(rays)),
        "sample": rays[::STRIDE].copy(),
        "sha256": np.frombuffer(hashlib.sha256(rays.tobytes()).digest(), dtype=np.uint8),
        "sum": rays.astype(np.float64).sum(axis=0),
        "abs_sum": np.abs(rays.astype(np.float64)).sum(axis=0),
    }


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    np.random.seed(12345)
    # ---- poses ----
    yaws = np.array([0.0, 0.3, -1.1, 2.5, np.pi])
    xyz = np.array([[0, 0, 0], [3.137, 2.718, 1.0], [-4.5, 7.25, 0.5], [19.0, 1.25, 1.0], [1e-3, -2e3, 3.0]], float)
    save("poses.npz", xyzyaw=np.concatenate([xyz, yaws[:, None]], 1),
         pose=np.stack([pose_of(*xyz[i], yaws[i]) for i in range(len(yaws))]))

    # ---- single-axis tables ----
    presets = {
        "8line": Indoor8LineLidarIntrinsics.create_standard_8line(),
        "32line": Indoor8LineLidarIntrinsics.create_dense_32line(),
        "64line": Indoor8LineLidarIntrinsics.create_leica_blk2go(),
        "lowcost": Indoor8LineLidarIntrinsics.create_low_cost_8line(),
        "custom": Indoor8LineLidarIntrinsics.create_custom_lidar(num_beams=3, beam_angles=[22.5, -1.25, -40.0], horizontal_resolution=0.7),
    }
    out = {}
    for pname, intr in presets.items():
        out[f"{pname}/vertical_degrees"] = np.asarray(intr.vertical_degrees, float)
        out[f"{pname}/W"] = np.int64(intr.horizontal_res)
        for qname, pose in POSES.items():
            for k, v in table_record(IndoorLidar(intr, pose).get_rays()).items():
                out[f"{pname}/{qname}/{k}"] = v
    save("rays_single_axis.npz", **out)

    # ---- uniform-fov branch (vertical_degrees=None) ----
    out = {}
    for (H, W) in ((1, 16), (4, 50), (8, 360)):
        intr = Indoor8LineLidarIntrinsics(vertical_res=H, horizontal_res=W, vertical_degrees=None)
        for qname, pose in POSES.items():
            out[f"{H}x{W}/{qname}/rays"] = IndoorLidar(intr, pose).get_rays()
    save("rays_uniform.npz", **out)

    # ---- dual-axis, noise off ----
    out = {}
    intr = DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    intr.angle_noise_std = 0.0
    intr.dropout_probability = 0.0
    for qname, pose in POSES.items():
        for k, v in table_record(DualAxisLidar(intr, pose).get_rays()).items():
            out[f"blk2go/{qname}/{k}"] = v
    small = DualAxisLidarIntrinsics(point_rate=1000, scan_duration=0.5, num_vertical_lines=7, swing_frequency=3.0,
                                    swing_amplitude=0.3, angle_noise_std=0.0, dropout_probability=0.0)
    out["small/rays"] = DualAxisLidar(small, POSES["posed"]).get_rays()
    save("rays_dual_axis.npz", **out)

    # ---- the reference's engine code on top of the oracle intersector ----
    from lrc_b200 import synthetic
    mesh = synthetic.box_room(target_tris=6000, seed=3)
    legacy = _LegacyMesh(mesh.vertices, mesh.triangles)
    eng = ref_cpu.RaycastEngineCPU()
    out = {"verts": mesh.vertices.astype(np.float32), "tris": mesh.triangles, "labels": mesh.triangle_labels}
    lid8 = create_lidar(Indoor8LineLidarIntrinsics.create_standard_8line(), POSES["posed"])
    pts, inc = eng.lidar_intersect_mesh(lid8, legacy)
    out["8line/points"], out["8line/incident"] = pts, inc
    # a short-range sensor so that the strict '<' range filter actually removes points
    short = Indoor8LineLidarIntrinsics(max_range=3.0, horizontal_res=500)
    pts, inc = eng.lidar_intersect_mesh(create_lidar(short, POSES["posed"]), legacy)
    out["short/points"], out["short/incident"] = pts, inc
    # sensor outside the room looking away on half of the azimuths: misses exercise the compaction
    outside = create_lidar(Indoor8LineLidarIntrinsics(horizontal_res=400), pose_of(-3.0, 4.0, 1.0, 0.0))
    rays = outside.get_rays()
    out["outside/rays"] = rays
    out["outside/points"] = eng.rays_intersect_mesh(rays=rays, mesh=legacy)
    pts, inc = eng.lidar_intersect_mesh(outside, legacy)
    out["outside/lidar_points"], out["outside/lidar_incident"] = pts, inc
    save("frame_box_room.npz", **out)


if __name__ == "__main__":
    main()
