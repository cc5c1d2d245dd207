"""
Posed sensors: a set of intrinsics at a 4x4 world pose, able to enumerate their rays.

API mirror of the reference's ``lidar/indoor_lidar.py`` (IndoorLidar :11-144, DualAxisLidar :146-374,
create_lidar :377-393, get_lidar_type :396-414).  ``get_rays()`` -- the only method the simulator's hot
loop reaches (reference raycast_engine_cpu.py:91) -- runs on the GPU through ``lrc_gen_rays_*``; there is
no host implementation of it.  When a posed sensor is handed to ``RaycastEngineGPU.lidar_intersect_mesh``
the rays are not materialised at all: they are generated inside the traversal kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Union

import numpy as np

from .intrinsics import DualAxisLidarIntrinsics, Indoor8LineLidarIntrinsics


def _check_pose(pose) -> None:
    assert isinstance(pose, np.ndarray)
    assert pose.shape == (4, 4)


@dataclass
class IndoorLidar:
    """Single-axis multi-line LiDAR at a pose (reference indoor_lidar.py:11-25)."""

    intrinsics: Indoor8LineLidarIntrinsics
    pose: np.ndarray

    def __post_init__(self):
        assert isinstance(self.intrinsics, Indoor8LineLidarIntrinsics)
        _check_pose(self.pose)

    def get_rays(self) -> np.ndarray:
        """(H*W, 6) float32 world rays [origin | unit direction], index = line*W + azimuth
        (reference indoor_lidar.py:27-53,94-131; uniform-fov variant :56-91 when vertical_degrees is None)."""
        from ..core import get_context
        rays, _ = get_context().gen_rays(self.pose[None], self.intrinsics)
        return rays.cpu().numpy()

    def get_total_rays(self) -> int:
        from ..core import rays_per_frame
        return rays_per_frame(self.intrinsics)

    def get_scan_frequency(self) -> float:
        return self.intrinsics.get_scan_frequency()

    def get_range_limits(self) -> tuple:
        return self.intrinsics.get_range_limits()


@dataclass
class DualAxisLidar:
    """Dual-axis scanner at a pose (reference indoor_lidar.py:146-160).

    ``seed`` / ``frame_index`` select the Philox stream of the angle noise and dropout.  The reference
    draws from numpy's global generator (indoor_lidar.py:271-272,293); to keep its "fresh noise on every
    call unless np.random.seed() was called" behaviour, an unset seed is itself drawn from that generator.
    """

    intrinsics: DualAxisLidarIntrinsics
    pose: np.ndarray
    seed: Optional[int] = None
    frame_index: int = 0

    def __post_init__(self):
        assert isinstance(self.intrinsics, DualAxisLidarIntrinsics)
        _check_pose(self.pose)

    # ---- hot path -------------------------------------------------------------------------------
    def resolve_seed(self) -> int:
        if self.seed is None:
            self.seed = int(np.random.randint(0, 2 ** 62, dtype=np.int64))
        return self.seed

    def noise_config(self):
        from ..core import NoiseConfig
        return NoiseConfig.from_intrinsics(self.intrinsics, seed=self.resolve_seed(), pose_index_base=self.frame_index)

    def get_multi_line_rays(self, num_points: int = None) -> np.ndarray:
        """Rays of the swinging-line pattern, dropped rays removed (reference indoor_lidar.py:224-296)."""
        from ..core import get_context
        intr = self.intrinsics
        if num_points is not None:
            # the reference sizes the table as num_points // num_vertical_lines per line (:241-244)
            from dataclasses import replace
            intr = replace(intr, point_rate=int(num_points), scan_duration=1.0)
        rays, keep = get_context().gen_rays(self.pose[None], intr, self.noise_config())
        rays = rays.cpu().numpy()
        return rays[keep.cpu().numpy().astype(bool)]

    def get_rays(self) -> np.ndarray:
        return self.get_multi_line_rays()                       # reference indoor_lidar.py:311-319

    def get_total_rays(self) -> int:
        return int(self.intrinsics.point_rate * self.intrinsics.scan_duration)

    def get_scan_frequency(self) -> float:
        return 1.0 / self.intrinsics.scan_duration

    def get_range_limits(self) -> tuple:
        return (0.5, self.intrinsics.max_range)

    # ---- time-parameterised helpers: NOT on the simulator's path (never called by the reference's
    # run_simulation); small host-side numpy utilities kept for API completeness -------------------
    def _direction_at(self, t: float) -> np.ndarray:
        phi, theta = self.intrinsics.calculate_angles_at_time(t, line_idx=0)
        return np.array([np.cos(theta) * np.cos(phi), np.cos(theta) * np.sin(phi), np.sin(theta)])

    def get_rays_at_time(self, t: float) -> np.ndarray:
        d = (self.pose[:3, :3] @ self._direction_at(t).astype(np.float32)).astype(np.float32)
        return np.concatenate([self.pose[:3, 3].astype(np.float32), d]).reshape(1, 6)

    def get_rays_sequence(self, time_sequence: np.ndarray) -> np.ndarray:
        o = self.pose[:3, 3].astype(np.float32)
        rows = [np.concatenate([o, (self.pose[:3, :3] @ self._direction_at(t)).astype(np.float32)]) for t in time_sequence]
        return np.array(rows, dtype=np.float32).reshape(-1, 6)

    def get_rays_frame(self, frame_duration: float = None) -> np.ndarray:
        return self.get_rays_sequence(self.intrinsics.generate_time_sequence(frame_duration))

    def get_spiral_scan_rays(self, num_points: int = None):
        n = int(self.intrinsics.point_rate * self.intrinsics.scan_duration) if num_points is None else num_points
        stamps = np.linspace(0, self.intrinsics.scan_duration, n)
        return self.get_rays_sequence(stamps), stamps

    def add_noise_to_rays(self, rays: np.ndarray) -> np.ndarray:
        p = self.intrinsics.dropout_probability
        return rays[np.random.random(len(rays)) > p] if p > 0 else rays


LidarType = Union[IndoorLidar, DualAxisLidar]
IntrinsicsType = Union[Indoor8LineLidarIntrinsics, DualAxisLidarIntrinsics]


def create_lidar(intrinsics: IntrinsicsType, pose: np.ndarray) -> LidarType:
    """Intrinsics type -> posed sensor (reference indoor_lidar.py:377-393)."""
    if isinstance(intrinsics, DualAxisLidarIntrinsics):
        return DualAxisLidar(intrinsics=intrinsics, pose=pose)
    if isinstance(intrinsics, Indoor8LineLidarIntrinsics):
        return IndoorLidar(intrinsics=intrinsics, pose=pose)
    raise ValueError(f"Unsupported LiDAR intrinsics type: {type(intrinsics)}")


def get_lidar_type(intrinsics: IntrinsicsType) -> str:
    """Human-readable sensor family (reference indoor_lidar.py:396-414)."""
    if isinstance(intrinsics, DualAxisLidarIntrinsics):
        return "Dual-axis spiral scanning"
    if isinstance(intrinsics, Indoor8LineLidarIntrinsics):
        if getattr(intrinsics, "dual_axis", False):
            return "Single-axis simulated dual-axis"
        return f"{intrinsics.vertical_res}-line single-axis scanning"
    return "Unknown type"
