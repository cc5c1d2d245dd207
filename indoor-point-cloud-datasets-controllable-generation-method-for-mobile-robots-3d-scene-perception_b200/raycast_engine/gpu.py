"""
``RaycastEngineGPU`` -- the drop-in for the engine object ``S3DISSimulator`` holds
(reference s3dis_simulator.py:67-74) and calls once per waypoint (:261).

Same class name, constructor and two methods as the reference's
``raycast_engine/raycast_engine_gpu_simple.py:12-98`` (which, despite its name, runs Open3D on the CPU);
same input checks and error types as ``raycast_engine_cpu.py:40-43``.  Everything below the Python
surface is the CUDA library behind ``include/lrc.h``.  There is no CPU path: without a B200 and a built
``csrc/liblrc.so`` the constructor raises.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from ..core import (Context, NoiseConfig, ScanResult, get_context, is_dual_axis, mesh_arrays)
from ..lidar.sensors import DualAxisLidar, IndoorLidar
from .base import RaycastEngineBase


class RaycastEngineGPU(RaycastEngineBase):
    """B200 ray-casting engine with the reference engine's interface.

    Extras beyond the reference interface: ``simulate`` (a whole trajectory in one fused launch
    sequence, device tensors out), ``last_scan`` (triangle ids / labels / ray indices of the most recent
    frame) and ``cast_rays``.
    """

    def __init__(self, verbose=False, device: Optional[int] = None, cache_mesh: bool = True):
        super().__init__()
        self.verbose = verbose
        self.cache_mesh = cache_mesh
        self.ctx: Context = get_context(device)
        self._last_scan: Optional[ScanResult] = None
        self._last_frame = None           # (pose, intrinsics, noise) of the last per-frame call: ``last_scan`` on demand

    # ---- reference interface ------------------------------------------------------------------
    def rays_intersect_mesh(self, rays: np.ndarray, mesh):
        """Closest hit of every ray; returns the hit points of the rays that hit, in ray order
        (reference raycast_engine_cpu.py:24-73)."""
        if not isinstance(rays, np.ndarray):
            raise TypeError("rays must be a numpy array.")
        if rays.ndim != 2 or rays.shape[1] != 6:
            raise ValueError("rays must be a (N, 6) array.")
        self._prepare(mesh)
        res = self.ctx.rays_intersect(rays.astype(np.float32))
        self.last_scan = res
        return res.points.cpu().numpy()

    def lidar_intersect_mesh(self, lidar, mesh):
        """One LiDAR frame: rays -> closest hits -> range filter -> incident angles
        (reference raycast_engine_cpu.py:75-111)."""
        self._prepare(mesh)
        if isinstance(lidar, (IndoorLidar, DualAxisLidar)):
            # one frame per call is the reference's call pattern (s3dis_simulator.py:254-263): one library call, one
            # synchronisation, results through the context's page-locked staging.  Triangle ids / labels / ray indices
            # of the frame are produced only when somebody asks for ``last_scan``.
            noise = lidar.noise_config() if isinstance(lidar, DualAxisLidar) else None
            pose = np.array(lidar.pose, dtype=np.float64)
            self._last_scan, self._last_frame = None, (pose, lidar.intrinsics, noise, self.ctx.stat("mesh_generation"))
            return self.ctx.scan_frame_to_host(pose, lidar.intrinsics, noise)     # empty frame -> np.empty(0), reference :109
        else:
            # duck-typed sensor (the reference touches only get_rays(), pose[:3,3], intrinsics.max_range)
            rays = lidar.get_rays()
            if not isinstance(rays, np.ndarray):
                raise TypeError("rays must be a numpy array.")
            if rays.ndim != 2 or rays.shape[1] != 6:
                raise ValueError("rays must be a (N, 6) array.")
            res = self.ctx.scan_rays(rays.astype(np.float32), np.asarray(lidar.pose, dtype=np.float64)[:3, 3],
                                     float(lidar.intrinsics.max_range))
        self.last_scan = res
        return self.ctx.frame_to_numpy(res)                                          # empty frame -> np.empty(0), reference :109

    # ---- extensions ---------------------------------------------------------------------------
    @property
    def last_scan(self) -> Optional[ScanResult]:
        """Device-resident record of the most recent call (points, incident angles, triangle ids, labels, ray indices).
        After a per-frame ``lidar_intersect_mesh`` it is produced on first access by re-running that frame on the
        resident BVH (deterministic: same pose, same sensor, same noise counter)."""
        if self._last_scan is None and self._last_frame is not None:
            pose, intr, noise, generation = self._last_frame
            if self.ctx.stat("mesh_generation") != generation:
                raise RuntimeError("last_scan: the mesh of the last frame is no longer resident (another mesh was set since)")
            self._last_scan = self.ctx.scan(pose[None], intr, noise)
        return self._last_scan

    @last_scan.setter
    def last_scan(self, value) -> None:
        self._last_scan, self._last_frame = value, None

    def _prepare(self, mesh) -> None:
        built = self.ctx.set_mesh(mesh, cache=self.cache_mesh)
        if built and self.verbose:
            print(f"[RaycastEngineGPU] LBVH built: {self.ctx.bvh_info()}")

    def set_mesh(self, mesh) -> None:
        """Build the BVH of ``mesh`` now and pin the object: per-waypoint calls that pass this same object skip the
        content check (a CRC over all vertices, indices and labels, ~6 ms per million triangles) that otherwise guards
        the BVH cache.  Do not edit the mesh arrays in place while it is pinned; ``invalidate_mesh()`` un-pins."""
        built = self.ctx.pin_mesh(mesh)
        if built and self.verbose:
            print(f"[RaycastEngineGPU] LBVH built: {self.ctx.bvh_info()}")

    def invalidate_mesh(self) -> None:
        """Forget the pinned mesh and the cached fingerprint: the next call rebuilds the BVH."""
        self.ctx.invalidate_mesh()

    def cast_rays(self, rays: np.ndarray, mesh=None):
        """Dense closest-hit query: (t_hit float32 [N] (+inf = miss), primitive_ids uint32 [N] (0xFFFFFFFF = miss))
        -- the two ``RaycastingScene.cast_rays`` outputs the path uses (reference raycast_engine_cpu.py:51-53)."""
        if mesh is not None:
            self._prepare(mesh)
        t, pid = self.ctx.cast_rays(np.ascontiguousarray(rays, dtype=np.float32))
        return t.cpu().numpy(), pid.cpu().numpy().view(np.uint32)

    def simulate_to_host(self, poses, intrinsics, mesh=None, noise: Optional[NoiseConfig] = None, host=None,
                         chunk_poses: Optional[int] = None) -> dict:
        """``simulate`` returning numpy arrays (points, incident, label, frame_offset) in pinned host memory; the
        device-to-host copies are pipelined behind the traversal of later pose chunks (see Context.scan_to_host)."""
        if mesh is not None:
            self._prepare(mesh)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 4, 4)
        return self.ctx.scan_to_host(poses, intrinsics, noise, host=host, chunk_poses=chunk_poses)

    def simulate(self, poses, intrinsics, mesh=None, noise: Optional[NoiseConfig] = None) -> ScanResult:
        """All frames of a trajectory at once.  ``poses``: (P,4,4) float64 (``Waypoint.to_pose_matrix``).
        Frame p of the result equals ``lidar_intersect_mesh(create_lidar(intrinsics, poses[p]), mesh)``."""
        if mesh is not None:
            self._prepare(mesh)
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 4, 4)
        res = self.ctx.scan(poses, intrinsics, noise)
        self.last_scan = res
        return res
