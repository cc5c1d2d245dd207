"""
Furniture records and the planner's furniture collision checks -- the API of the reference's
``trajectory/collision_detector.py`` (``FurnitureInfo`` :13-41, ``CollisionDetector`` :44-262) behind
``AutoTrajectoryGenerator.add_furniture / add_furniture_from_mesh / clear_furniture``
(reference trajectory/auto_trajectory_generator.py:693-704).

The reference only STORES the furniture (its planner never queries the detector), and its per-waypoint test raises
``AttributeError`` for the first piece of furniture whose expanded box does not contain the robot (``furniture.mesh``
does not exist, :125).  Here the same box test is evaluated for all waypoints x all pieces at once with numpy and a
piece without a mesh simply is not refined -- the only deviation, documented in DESIGN.md.  Host-side: a few boxes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .trajectory_generator import Waypoint

_B_KEYS = ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")


@dataclass
class FurnitureInfo:
    """Position and size of a piece of furniture (the mesh itself occludes the LiDAR; this record is for planning)."""
    name: str
    position: np.ndarray          # (3,) centre
    size: np.ndarray              # (3,) extents [length, width, height]
    category: str = "unknown"

    def _lo_hi(self) -> Tuple[np.ndarray, np.ndarray]:
        c, h = np.asarray(self.position, dtype=np.float64), np.asarray(self.size, dtype=np.float64) / 2
        return c - h, c + h

    def get_bounds(self) -> Dict[str, float]:
        lo, hi = self._lo_hi()
        return {"x_min": lo[0], "x_max": hi[0], "y_min": lo[1], "y_max": hi[1], "z_min": lo[2], "z_max": hi[2]}

    def is_point_inside(self, point) -> bool:
        lo, hi = self._lo_hi()
        p = np.asarray(point, dtype=np.float64)
        return bool(np.all((lo <= p) & (p <= hi)))


class CollisionDetector:
    def __init__(self, robot_radius: float = 0.3):
        self.robot_radius = robot_radius
        self.furniture_list: List[FurnitureInfo] = []

    # ---- furniture list ----
    def add_furniture(self, furniture: FurnitureInfo) -> None:
        self.furniture_list.append(furniture)

    def add_furniture_from_mesh(self, mesh, name: str, category: str = "unknown") -> None:
        """Centre = mean of the vertices, size = their extent (reference :58-78); an empty mesh adds nothing."""
        v = np.asarray(mesh.vertices if hasattr(mesh, "vertices") else mesh[0], dtype=np.float64).reshape(-1, 3)
        if len(v) == 0:
            return
        self.add_furniture(FurnitureInfo(name=name, position=v.mean(axis=0), size=v.max(axis=0) - v.min(axis=0), category=category))

    def get_furniture_list(self) -> List[FurnitureInfo]:
        return list(self.furniture_list)

    def clear_furniture(self) -> None:
        self.furniture_list.clear()

    # ---- collision tests ----
    def _hits(self, positions: np.ndarray) -> np.ndarray:
        """(W, F) bool: robot centre inside furniture box grown by the robot radius on every side (closed intervals)."""
        if not self.furniture_list:
            return np.zeros((len(positions), 0), dtype=bool)
        lo = np.stack([f._lo_hi()[0] for f in self.furniture_list]) - self.robot_radius
        hi = np.stack([f._lo_hi()[1] for f in self.furniture_list]) + self.robot_radius
        p = positions[:, None, :]
        return np.all((lo[None] <= p) & (p <= hi[None]), axis=2)

    def detect_collision(self, waypoint: Waypoint) -> Tuple[bool, Optional[FurnitureInfo]]:
        """(collides, first colliding piece in list order)."""
        h = self._hits(np.array([[waypoint.x, waypoint.y, waypoint.z]], dtype=np.float64))[0]
        k = int(np.argmax(h)) if h.any() else -1
        return (True, self.furniture_list[k]) if k >= 0 else (False, None)

    def detect_path_collision(self, waypoints: Sequence[Waypoint]) -> List[Tuple[int, FurnitureInfo]]:
        if not waypoints:
            return []
        h = self._hits(np.array([[w.x, w.y, w.z] for w in waypoints], dtype=np.float64))
        return [(int(i), self.furniture_list[int(np.argmax(h[i]))]) for i in np.nonzero(h.any(axis=1))[0]]

    def suggest_avoidance_path(self, waypoint: Waypoint, collided_furniture: FurnitureInfo) -> List[Waypoint]:
        """Left bypass, right bypass, step back -- each robot_radius + 0.5 m away from the waypoint (reference :169-223)."""
        pos = np.array([waypoint.x, waypoint.y, waypoint.z], dtype=np.float64)
        to_f = np.asarray(collided_furniture.position, dtype=np.float64) - pos
        to_f[2] = 0.0
        n = np.linalg.norm(to_f)
        if n > 0:
            to_f = to_f / n
        dist = self.robot_radius + 0.5
        out = []
        for off in (-np.pi / 2, np.pi / 2):
            c, s = np.cos(off), np.sin(off)
            d = np.array([c * to_f[0] - s * to_f[1], s * to_f[0] + c * to_f[1], to_f[2]])
            q = pos + d * dist
            out.append(Waypoint(x=q[0], y=q[1], z=q[2], yaw=waypoint.yaw + off))
        q = pos - to_f * dist
        out.append(Waypoint(x=q[0], y=q[1], z=q[2], yaw=waypoint.yaw))
        return out

    def get_collision_statistics(self, waypoints: Sequence[Waypoint]) -> Dict[str, Any]:
        col = self.detect_path_collision(waypoints)
        per: Dict[str, int] = {}
        for _, f in col:
            per[f.name] = per.get(f.name, 0) + 1
        return {"total_collisions": len(col), "collision_rate": len(col) / len(waypoints) if waypoints else 0, "collision_furniture": per}
