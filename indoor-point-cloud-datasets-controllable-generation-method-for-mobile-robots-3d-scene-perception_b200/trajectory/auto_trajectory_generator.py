"""
``AutoTrajectoryGenerator`` -- the caller that produces the poses the ray-cast engine consumes
(reference trajectory/auto_trajectory_generator.py, used by s3dis_simulator.py:161).

Same class name, constructor arguments, tunables, ``generate_optimal_trajectory(mesh, room_bounds, num_waypoints)``
signature and ``analysis_info`` keys as the reference.  What differs is where the work happens:

* the occupancy test "any mesh vertex inside the robot cube" (reference :220-238), evaluated by the reference with a
  numpy pass over ALL vertices for every grid point and every waypoint of every candidate, runs on the GPU against a
  binned vertex index (``lrc_collision_index_build`` / ``lrc_collision_query``) -- one launch for the whole grid, one
  for all candidates' waypoints;
* the O(n^2) Python connectivity loop (:245-258) is ``lrc_grid_connectivity`` (CSR, same neighbour order);
* A* (:413-473, O(n^2) with a Python set) is a binary-heap search in the native library (``lrc_astar``).

Everything that decides a boolean (bounds tests, collisions, neighbour distances) uses the reference's float64
arithmetic and is bit-exact against it; with the same ``np.random`` state the same start / end candidates are drawn.
The one documented difference: among shortest paths of exactly equal cost the reference's choice follows CPython's set
iteration order, this planner's the smaller node index -- path cost is identical, path shape may differ.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _native as nat
from ..core import get_context
from .trajectory_generator import TrajectoryQuality, Waypoint

ROBOT_HEIGHT = 1.0                      # reference :122 ("Fixed robot height")
_B_KEYS = ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")


@dataclass
class RoomAnalysis:
    """Same fields as the reference's ``RoomAnalysis`` (:19-27); the graph is CSR instead of a dict of lists."""
    bounds: Dict[str, float]
    center: np.ndarray
    dimensions: np.ndarray
    free_space_points: np.ndarray        # (n,3) float64, reference order (x-major)
    obstacle_points: np.ndarray          # (m,3)
    connectivity_graph: Tuple[np.ndarray, np.ndarray]   # (row_ptr int32 [n+1], col int32 [nnz]), ascending columns
    mesh: object
    resolution: float = 0.0
    detailed: bool = False

    def neighbours(self, i: int) -> np.ndarray:
        ptr, col = self.connectivity_graph
        return col[ptr[i]:ptr[i + 1]]


@dataclass
class TrajectoryCandidate:
    """Same fields as the reference's ``TrajectoryCandidate`` (:30-39)."""
    start_point: np.ndarray
    end_point: np.ndarray
    waypoints: List[Waypoint]
    quality: TrajectoryQuality
    length: float
    collision_count: int
    smoothness_score: float


class AutoTrajectoryGenerator:
    def __init__(self, robot_radius: float = 0.3, min_trajectory_length: Optional[float] = None, device: Optional[int] = None):
        self.robot_radius = robot_radius
        self.min_trajectory_length = min_trajectory_length
        self.room_analysis: Optional[RoomAnalysis] = None
        # reference :53-61
        self.grid_resolution = 0.2
        self.min_free_space = 1.0
        self.max_candidates = 40
        self.sampling_density = 0.1
        self.interpolation_density = 2.0
        self.min_waypoints = 40
        self._device = device
        self._ctx = None
        from .collision_detector import CollisionDetector
        self.collision_detector = CollisionDetector(robot_radius)          # reference :51

    # ---- GPU plumbing ---------------------------------------------------------------------------------------
    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context(self._device)
        return self._ctx

    def _index_mesh(self, mesh) -> None:
        """Bin the mesh's float64 vertices once (cell = robot cube edge)."""
        v = np.ascontiguousarray(np.asarray(mesh.vertices if hasattr(mesh, "vertices") else mesh[0]), dtype=np.float64).reshape(-1, 3)
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            self._verts_d = torch.from_numpy(v).to(ctx.device)
            self._build_index()

    def _build_index(self) -> None:
        ctx = self.ctx
        nat.check(ctx._h, ctx._lib.lrc_collision_index_build(ctx._h, C.c_void_p(self._verts_d.data_ptr()), self._verts_d.shape[0],
                                                             max(2.0 * self.robot_radius, 1e-3), ctx._stream()))
        # the context holds ONE collision index: remember which build is ours, so a query after somebody else's build
        # (another planner on the same GPU) re-bins our vertices instead of silently testing theirs
        self._index_generation = ctx.stat("collision_generation")

    def _query(self, points: np.ndarray, bounds: Optional[Dict[str, float]]) -> np.ndarray:
        """state per point: 0 = robot cube leaves the room, 1 = a vertex inside the cube, 2 = free."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        if len(pts) == 0:
            return np.zeros(0, np.uint8)
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            if getattr(self, "_verts_d", None) is None:
                raise RuntimeError("no mesh indexed yet: call _index_mesh / generate_optimal_trajectory first")
            if ctx.stat("collision_generation") != self._index_generation:
                self._build_index()
            p_d = torch.from_numpy(pts).to(ctx.device)
            st = torch.empty(len(pts), dtype=torch.uint8, device=ctx.device)
            b = None if bounds is None else (C.c_double * 6)(*[float(bounds[k]) for k in _B_KEYS])
            nat.check(ctx._h, ctx._lib.lrc_collision_query(ctx._h, C.c_void_p(p_d.data_ptr()), len(pts), float(self.robot_radius), b,
                                                           C.c_void_p(st.data_ptr()), ctx._stream()))
            return st.cpu().numpy()

    def _grid(self, room_bounds, resolution: float):
        """Classify the grid samples of reference :123-146 and build the connectivity graph of the free ones."""
        xs = np.arange(room_bounds["x_min"], room_bounds["x_max"], resolution)
        ys = np.arange(room_bounds["y_min"], room_bounds["y_max"], resolution)
        nx, ny = len(xs), len(ys)
        gx, gy = np.meshgrid(xs, ys, indexing="ij")
        pts = np.stack([gx.ravel(), gy.ravel(), np.full(nx * ny, ROBOT_HEIGHT)], axis=1)
        state = self._query(pts, room_bounds)
        free, obstacles = pts[state == 2], pts[state == 1]
        ctx = self.ctx
        n = nx * ny
        ptr = np.zeros(1, np.int32)
        col = np.zeros(0, np.int32)
        if len(free) > 0:
            max_dist = self.robot_radius * 2                      # reference :248
            step = float(min(np.min(np.diff(xs)) if nx > 1 else resolution, np.min(np.diff(ys)) if ny > 1 else resolution))
            window = int(np.ceil(max_dist / step)) + 1
            with torch.cuda.device(ctx.device):
                dev = ctx.device
                xs_d, ys_d = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
                st_d = torch.from_numpy(state).to(dev)
                fi = torch.empty(n, dtype=torch.int32, device=dev)
                rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
                cap = len(free) * min((2 * window + 1) ** 2 - 1, max(len(free) - 1, 1))
                cl = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
                counts = (C.c_int64 * 2)()
                nat.check(ctx._h, ctx._lib.lrc_grid_connectivity(
                    ctx._h, C.c_void_p(xs_d.data_ptr()), nx, C.c_void_p(ys_d.data_ptr()), ny, C.c_void_p(st_d.data_ptr()),
                    float(max_dist), window, C.c_void_p(fi.data_ptr()), C.c_void_p(rp.data_ptr()), C.c_void_p(cl.data_ptr()),
                    int(cl.numel()), counts, ctx._stream()))
                nf, nnz = int(counts[0]), int(counts[1])
                assert nf == len(free)
                ptr = rp[:nf + 1].cpu().numpy()
                col = cl[:nnz].cpu().numpy()
        return free, obstacles, (ptr, col)

    # ---- reference API --------------------------------------------------------------------------------------
    def generate_optimal_trajectory(self, mesh, room_bounds: Dict[str, float], num_waypoints: int = 20):
        """-> (waypoints, analysis_info); reference :64-95."""
        self.room_analysis = self._analyze_room_layout(mesh, room_bounds)
        dense_waypoints = max(int(num_waypoints * self.interpolation_density), self.min_waypoints)
        candidates = self._generate_trajectory_candidates(dense_waypoints)
        best = self._select_best_trajectory(candidates)
        return best.waypoints, self._generate_analysis_info(candidates, best)

    def _analyze_room_layout(self, mesh, room_bounds: Dict[str, float]) -> RoomAnalysis:
        """Reference :97-202: coarse grid, detailed grid when fewer than 10 free samples."""
        center = np.array([(room_bounds["x_max"] + room_bounds["x_min"]) / 2, (room_bounds["y_max"] + room_bounds["y_min"]) / 2,
                           (room_bounds["z_max"] + room_bounds["z_min"]) / 2])
        dimensions = np.array([room_bounds["x_max"] - room_bounds["x_min"], room_bounds["y_max"] - room_bounds["y_min"],
                               room_bounds["z_max"] - room_bounds["z_min"]])
        if self.min_trajectory_length is None:
            self.min_trajectory_length = max(dimensions[0], dimensions[1]) * 0.2
        self._index_mesh(mesh)
        resolution = max(0.2, min(dimensions) / 20)
        free, obstacles, graph = self._grid(room_bounds, resolution)
        detailed = False
        if len(free) < 10:
            detailed = True
            resolution = max(0.15, min(dimensions) / 30)
            free, obstacles, graph = self._grid(room_bounds, resolution)
        return RoomAnalysis(bounds=room_bounds, center=center, dimensions=dimensions, free_space_points=free,
                            obstacle_points=obstacles, connectivity_graph=graph, mesh=mesh, resolution=resolution, detailed=detailed)

    def _a_star_search(self, start_idx: int, end_idx: int) -> Optional[List[int]]:
        ra = self.room_analysis
        ptr, col = ra.connectivity_graph
        free = np.ascontiguousarray(ra.free_space_points, dtype=np.float64)
        n = len(free)
        path = np.empty(n, np.int32)
        ln, cost = C.c_int32(0), C.c_double(0.0)
        ptr = np.ascontiguousarray(ptr, np.int32)
        col = np.ascontiguousarray(col, np.int32)
        nat.check(None, nat.load().lrc_astar(C.c_void_p(ptr.ctypes.data), C.c_void_p(col.ctypes.data if len(col) else 0),
                                             C.c_void_p(free.ctypes.data), n, int(start_idx), int(end_idx),
                                             C.c_void_p(path.ctypes.data), n, C.byref(ln), C.byref(cost)))
        if ln.value == 0:
            return None
        self._last_path_cost = cost.value
        return [int(i) for i in path[:ln.value]]

    def _generate_trajectory_candidates(self, num_waypoints: int) -> List[TrajectoryCandidate]:
        """Reference :260-298 (same draws from the global numpy stream) + :300-384, with the collision counts of all
        candidates evaluated in one GPU query."""
        free = self.room_analysis.free_space_points
        if len(free) < 2:
            return []
        max_attempts = min(self.max_candidates, len(free) * 2)
        drafts = []
        for _ in range(max_attempts):
            start_idx = np.random.randint(0, len(free))
            end_idx = np.random.randint(0, len(free))
            if start_idx == end_idx:
                continue
            start_point, end_point = free[start_idx], free[end_idx]
            if np.linalg.norm(start_point - end_point) < self.min_trajectory_length:
                continue
            w = self._candidate_polyline(start_idx, end_idx, num_waypoints)
            if w is not None and len(w) > 0:
                drafts.append((start_point, end_point, w))
        if not drafts:
            return []
        # collisions (:352-362): outside the room bounds OR a vertex inside the robot cube
        state = self._query(np.concatenate([w for _, _, w in drafts]), self.room_analysis.bounds)
        out, at = [], 0
        for start_point, end_point, w in drafts:
            st = state[at:at + len(w)]
            at += len(w)
            collision_count = int(np.count_nonzero(st != 2))
            wps = [Waypoint(x=p[0], y=p[1], z=p[2], yaw=0) for p in w]
            length = self._calculate_trajectory_length(w)
            smoothness = self._calculate_smoothness_score(wps)
            quality = TrajectoryQuality(
                coverage_ratio=1.0 - (collision_count / len(wps)), path_length=length, turn_count=self._count_turns(w),
                efficiency=1.0 if collision_count == 0 else max(0.0, 1.0 - collision_count / len(wps)),
                collision_count=collision_count, smoothness=smoothness)
            out.append(TrajectoryCandidate(start_point, end_point, wps, quality, length, collision_count, smoothness))
        return out

    def _candidate_polyline(self, start_idx: int, end_idx: int, num_waypoints: int) -> Optional[np.ndarray]:
        """(m,3) waypoint positions of one candidate: A* over the free-space graph, resampled and smoothed (:300-348)."""
        free = self.room_analysis.free_space_points
        start_point, end_point = free[start_idx], free[end_idx]
        # the nearest free sample of a free sample is itself (:314-315)
        path = self._a_star_search(start_idx, end_idx)
        if path is None or len(path) < 2:
            return self._generate_linear_waypoints(start_point, end_point, num_waypoints)
        path_points = [free[i] for i in path]
        if len(path_points) == 2:
            return self._generate_linear_waypoints(path_points[0], path_points[1], num_waypoints)
        return self._smooth_trajectory(self._generate_waypoints_along_path(path_points, num_waypoints))

    # ---- polyline helpers: same arithmetic as the reference, on (n,3) arrays --------------------------------------
    @staticmethod
    def _generate_linear_waypoints(a, b, n: int) -> np.ndarray:
        """Reference :386-398."""
        out = np.empty((n, 3))
        for i in range(n):
            t = i / (n - 1) if n > 1 else 0
            out[i] = (a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]), a[2] + t * (b[2] - a[2]))
        return out

    @staticmethod
    def _generate_waypoints_along_path(path_points, n: int) -> np.ndarray:
        """Reference :475-527: ``n`` samples at equal arc length, the last one pinned to the path's end."""
        pp = np.asarray(path_points, dtype=np.float64)
        if len(pp) < 2:
            return np.zeros((0, 3))
        seg = [np.linalg.norm(pp[i + 1] - pp[i]) for i in range(len(pp) - 1)]
        total = 0.0
        for s in seg:
            total += s
        if total < 1e-6:
            return pp[:1].copy()
        ends = []
        acc = 0.0
        for s in seg:
            acc = acc + s
            ends.append(acc)
        out = []
        for i in range(n):
            if i == n - 1:
                out.append(pp[-1])
                break
            target = (i / (n - 1)) * total
            for k, s in enumerate(seg):
                if target <= ends[k]:
                    seg_start = ends[k - 1] if k > 0 else 0.0
                    prog = (target - seg_start) / s if s > 0 else 0
                    out.append(pp[k] + prog * (pp[k + 1] - pp[k]))
                    break
        return np.array(out)

    @staticmethod
    def _smooth_trajectory(w: np.ndarray, alpha: float = 0.5) -> np.ndarray:
        """Reference :529-554: interior points pulled halfway towards the mean of their ORIGINAL neighbours."""
        w = np.asarray(w, dtype=np.float64)
        if len(w) < 3:
            return w
        out = w.copy()
        out[1:-1] = alpha * w[1:-1] + (1 - alpha) * (w[:-2] + w[2:]) / 2
        return out

    @staticmethod
    def _count_turns(w: np.ndarray) -> int:
        """Reference :556-593: direction changes above 30 degrees."""
        if len(w) < 3:
            return 0
        turns = 0
        for i in range(1, len(w) - 1):
            v1, v2 = w[i, :2] - w[i - 1, :2], w[i + 1, :2] - w[i, :2]
            n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
            if n1 > 1e-6 and n2 > 1e-6:
                if np.arccos(np.clip(np.dot(v1 / n1, v2 / n2), -1.0, 1.0)) > np.pi / 6:
                    turns += 1
        return turns

    @staticmethod
    def _calculate_trajectory_length(w: np.ndarray) -> float:
        """Reference :595-612 (sequential float64 sum)."""
        total = 0.0
        for i in range(1, len(w)):
            total += np.sqrt((w[i, 0] - w[i - 1, 0]) ** 2 + (w[i, 1] - w[i - 1, 1]) ** 2 + (w[i, 2] - w[i - 1, 2]) ** 2)
        return total

    @staticmethod
    def _calculate_smoothness_score(waypoints: List[Waypoint]) -> float:
        """Reference :614-635: 1 - std(|yaw changes|) / pi (all yaws are 0 on this path, so 1.0)."""
        if len(waypoints) < 3:
            return 1.0
        changes = [abs(waypoints[i].yaw - waypoints[i - 1].yaw) for i in range(1, len(waypoints))]
        return max(0, 1 - np.std(changes) / np.pi)

    def _select_best_trajectory(self, candidates: List[TrajectoryCandidate]) -> TrajectoryCandidate:
        """Reference :637-665: 0.4 * min(len / min_len, 2) + 0.4 * smoothness - 0.1 * collisions; first best wins."""
        if not candidates:
            raise ValueError("No available trajectory candidates")
        best, best_score = None, -1
        for c in candidates:
            score = min(c.length / self.min_trajectory_length, 2.0) * 0.4 + c.smoothness_score * 0.4 - c.collision_count * 0.1
            if score > best_score:
                best, best_score = c, score
        return best

    def _generate_analysis_info(self, candidates, best) -> Dict[str, Any]:
        """Reference :667-704, same keys."""
        if not candidates:
            return {}
        lengths = [c.length for c in candidates]
        collisions = [c.collision_count for c in candidates]
        smooth = [c.smoothness_score for c in candidates]
        ra = self.room_analysis
        return {
            "total_candidates": len(candidates),
            "best_trajectory": {"length": best.length, "collision_count": best.collision_count,
                                "smoothness_score": best.smoothness_score, "start_point": best.start_point.tolist(),
                                "end_point": best.end_point.tolist()},
            "statistics": {"length_mean": np.mean(lengths), "length_std": np.std(lengths), "collision_mean": np.mean(collisions),
                           "collision_std": np.std(collisions), "smoothness_mean": np.mean(smooth), "smoothness_std": np.std(smooth)},
            "room_analysis": {"free_space_points": len(ra.free_space_points), "obstacle_points": len(ra.obstacle_points),
                              "room_dimensions": ra.dimensions.tolist(), "room_center": ra.center.tolist()},
        }

    # ---- furniture records (reference :693-704; the reference stores them and never queries them while planning) ----
    def add_furniture(self, furniture) -> None:
        self.collision_detector.add_furniture(furniture)

    def add_furniture_from_mesh(self, mesh, name: str, category: str = "unknown") -> None:
        self.collision_detector.add_furniture_from_mesh(mesh, name, category)

    def clear_furniture(self) -> None:
        self.collision_detector.clear_furniture()
