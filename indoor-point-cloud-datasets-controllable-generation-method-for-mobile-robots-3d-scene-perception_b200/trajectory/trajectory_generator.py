"""
Pose format of the simulator's frame loop (reference trajectory/trajectory_generator.py:13-44 and
s3dis_simulator.py:254-257) plus simple seeded trajectories for the synthetic benchmark scenes.
The coverage planner lives next door in ``auto_trajectory_generator.py``.
"""
from __future__ import annotations

from dataclasses import dataclass
from dataclasses import asdict
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np


@dataclass
class Waypoint:
    """Position + yaw (radians) of the sensor (reference trajectory_generator.py:13-28)."""
    x: float
    y: float
    z: float
    yaw: float
    timestamp: float = 0.0
    velocity: Optional[float] = None
    angular_velocity: Optional[float] = None

    def to_array(self) -> np.ndarray:
        return np.array([self.x, self.y, self.z, self.yaw])

    def to_pose_matrix(self) -> np.ndarray:
        """4x4 float64: translation (x,y,z), rotation about +Z by yaw (reference :30-44)."""
        c, s = np.cos(self.yaw), np.sin(self.yaw)
        m = np.eye(4)
        m[:3, 3] = (self.x, self.y, self.z)
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
        return m

    def distance_to(self, other: "Waypoint") -> float:
        return float(np.sqrt((self.x - other.x) ** 2 + (self.y - other.y) ** 2 + (self.z - other.z) ** 2))

    def angle_to(self, other: "Waypoint") -> float:
        return float(np.arctan2(other.y - self.y, other.x - self.x))


@dataclass
class TrajectoryQuality:
    """Same fields as the reference's ``TrajectoryQuality`` (trajectory_generator.py:60-82)."""
    coverage_ratio: float
    path_length: float
    turn_count: int
    efficiency: float
    collision_count: int
    smoothness: float

    def to_dict(self) -> Dict[str, Any]:
        return asdict(self)


def poses_from_waypoints(waypoints: Iterable[Waypoint]) -> np.ndarray:
    """(P,4,4) float64 stack of ``to_pose_matrix()`` -- what ``RaycastEngineGPU.simulate`` consumes."""
    mats = [w.to_pose_matrix() for w in waypoints]
    return np.stack(mats) if mats else np.zeros((0, 4, 4))


def polyline_waypoints(vertices: Sequence[Sequence[float]], count: int, z: float = 1.0, yaw: float = 0.0) -> List[Waypoint]:
    """``count`` waypoints at equal arc-length spacing along a 2-D polyline.  z = 1.0 and yaw = 0 are what the
    reference's auto generator emits (auto_trajectory_generator.py:122,400)."""
    pts = np.asarray(vertices, dtype=np.float64)
    seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
    cum = np.concatenate([[0.0], np.cumsum(seg)])
    s = np.linspace(0.0, cum[-1], count) if count > 1 else np.array([0.0])
    out = []
    for k, d in enumerate(s):
        i = min(int(np.searchsorted(cum, d, side="right")) - 1, len(seg) - 1)
        f = 0.0 if seg[i] == 0 else (d - cum[i]) / seg[i]
        p = pts[i] + f * (pts[i + 1] - pts[i])
        out.append(Waypoint(float(p[0]), float(p[1]), float(z), float(yaw), timestamp=0.1 * k))
    return out


def shard_range(num_poses: int, rank: int, world: int) -> range:
    """Contiguous pose slice of rank ``rank``: [floor(r*P/G), floor((r+1)*P/G))."""
    return range(rank * num_poses // world, (rank + 1) * num_poses // world)
