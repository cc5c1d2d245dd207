"""
Golden fixtures for the coverage-trajectory planner, produced by the LIVE REFERENCE's own
trajectory/auto_trajectory_generator.py (imported with a stub ``open3d``; the planner only reads ``mesh.vertices``).
Run once in the build container:

    python tests/golden/make_golden_plan.py

  plan_room.npz    box room (~6k triangles, furniture): bounds, grid classification (free / obstacle points in the
                   reference's order), connectivity graph (CSR), A* paths + costs for seeded index pairs, the polyline
                   helpers on one of those paths, collision verdicts for seeded query points, and the trajectory
                   generate_optimal_trajectory returns under np.random.seed(7)
  plan_tight.npz   a small cluttered room whose coarse grid has < 10 free points -> the reference's detailed branch
The meshes are regenerated from lrc_b200.synthetic at test time (seeded), only the planner outputs are stored.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
o3d = types.ModuleType("open3d")
o3d.geometry = types.SimpleNamespace(TriangleMesh=object, PointCloud=object, AxisAlignedBoundingBox=object)
o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: a)
sys.modules["open3d"] = o3d
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
from trajectory.auto_trajectory_generator import AutoTrajectoryGenerator  # noqa: E402
from trajectory.trajectory_generator import Waypoint  # noqa: E402
import lrc_b200 as lrc  # noqa: E402


class _Mesh:
    def __init__(self, m):
        self.vertices, self.triangles = m.vertices, m.triangles


def bounds_of(v):
    return dict(x_min=float(v[:, 0].min()), x_max=float(v[:, 0].max()), y_min=float(v[:, 1].min()), y_max=float(v[:, 1].max()),
                z_min=float(v[:, 2].min()), z_max=float(v[:, 2].max()))


def csr(graph):
    n = len(graph)
    ptr = np.zeros(n + 1, np.int64)
    for i in range(n):
        ptr[i + 1] = ptr[i] + len(graph[i])
    col = np.array([j for i in range(n) for j in graph[i]], dtype=np.int32)
    return ptr, col


def capture(name, mesh, n_pairs, seed):
    m = _Mesh(mesh)
    b = bounds_of(mesh.vertices)
    g = AutoTrajectoryGenerator()
    ra = g._analyze_room_layout(m, b)
    g.room_analysis = ra
    free = np.array(ra.free_space_points).reshape(-1, 3)
    obst = np.array(ra.obstacle_points).reshape(-1, 3)
    ptr, col = csr(ra.connectivity_graph)
    out = {"bounds": np.array([b[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")]),
           "free": free, "obstacles": obst, "graph_ptr": ptr, "graph_col": col,
           "min_trajectory_length": np.float64(g.min_trajectory_length)}
    rng = np.random.default_rng(seed)
    pairs, costs, lens = [], [], []
    paths = []
    for _ in range(n_pairs):
        a, c = int(rng.integers(0, len(free))), int(rng.integers(0, len(free)))
        path = g._a_star_search(a, c, ra.free_space_points)
        pairs.append((a, c))
        if path is None:
            costs.append(-1.0)
            lens.append(0)
        else:
            costs.append(float(sum(np.linalg.norm(free[path[k + 1]] - free[path[k]]) for k in range(len(path) - 1))))
            lens.append(len(path))
            paths.append(path)
    out["pairs"], out["astar_cost"], out["astar_len"] = np.array(pairs, np.int32), np.array(costs), np.array(lens, np.int32)
    # polyline helpers on the longest path found
    path = max(paths, key=len)
    pp = [ra.free_space_points[i] for i in path]
    w = g._generate_waypoints_along_path(pp, 40)
    ws = g._smooth_trajectory(w)
    arr = lambda ws_: np.array([[q.x, q.y, q.z] for q in ws_])
    out["helper/path"] = np.array(path, np.int32)
    out["helper/along"], out["helper/smooth"] = arr(w), arr(ws)
    out["helper/turns"] = np.int64(g._count_turns(ws))
    out["helper/length"] = np.float64(g._calculate_trajectory_length(ws))
    out["helper/turns_raw"] = np.int64(g._count_turns(w))
    # collision verdicts for seeded query points (inside and outside the room)
    q = np.stack([rng.uniform(b["x_min"] - 0.5, b["x_max"] + 0.5, 300), rng.uniform(b["y_min"] - 0.5, b["y_max"] + 0.5, 300),
                  rng.uniform(0.2, 2.9, 300)], axis=1)
    out["query/points"] = q
    out["query/in_bounds"] = np.array([g._is_point_in_room_bounds(p, b) for p in q])
    out["query/collides"] = np.array([g._is_point_inside_mesh(p, m) for p in q])
    # the complete planner under a fixed seed
    np.random.seed(7)
    g2 = AutoTrajectoryGenerator()
    wps, info = g2.generate_optimal_trajectory(m, b, num_waypoints=20)
    out["traj/waypoints"] = np.array([[q_.x, q_.y, q_.z, q_.yaw] for q_ in wps])
    out["traj/length"] = np.float64(info["best_trajectory"]["length"])
    out["traj/collisions"] = np.int64(info["best_trajectory"]["collision_count"])
    out["traj/total_candidates"] = np.int64(info["total_candidates"])
    out["traj/start"], out["traj/end"] = np.array(info["best_trajectory"]["start_point"]), np.array(info["best_trajectory"]["end_point"])
    path_ = os.path.join(HERE, name)
    np.savez_compressed(path_, **out)
    print(name, os.path.getsize(path_), "bytes;", len(free), "free,", len(obst), "obstacle,", len(col), "edges; traj", len(wps), "waypoints, length",
          float(out["traj/length"]))


def main():
    capture("plan_room.npz", lrc.synthetic.box_room(target_tris=6000, seed=3), n_pairs=12, seed=11)
    capture("plan_tight.npz", lrc.synthetic.planner_tight_room(), n_pairs=4, seed=12)


if __name__ == "__main__":
    main()
