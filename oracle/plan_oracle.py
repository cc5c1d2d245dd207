"""
CPU ORACLE for the coverage-trajectory planner -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py for who may import it).

Restates /root/reference/trajectory/auto_trajectory_generator.py in numpy / pure Python, statement by statement where
the arithmetic decides a result (comparisons, distances, interpolation), vectorised only where that cannot change a bit:

  in_room_bounds        :204-217   robot cube [p - r, p + r] inside the room bounds (closed comparisons)
  collides              :220-238   "any mesh vertex inside the closed robot cube", float64 vertices as stored in the mesh
  analyze_room_layout   :97-202    grid sampling at z = 1.0 with np.arange, x-major / y-minor order, coarse -> detailed
  connectivity_graph    :245-258   j != i and ||p_i - p_j|| <= 2 r, neighbours in ascending j
  a_star                :413-473   open set as a Python set, min() by f-score (tie-break = set iteration order)
  waypoints_along_path  :475-527, smooth :529-554, count_turns :556-593, path_length :595-612, smoothness :614-635
  candidate scoring     :637-665   0.4 * min(len / min_len, 2) + 0.4 * smoothness - 0.1 * collisions, first best wins

Pinned by tests/golden/plan_*.npz, produced by the reference's own AutoTrajectoryGenerator
(tests/golden/make_golden_plan.py).
"""
from __future__ import annotations

import numpy as np

ROBOT_HEIGHT = 1.0          # auto_trajectory_generator.py:122


def in_room_bounds(points, bounds, r):
    """:204-217 for an (n,3) array; bounds = dict x_min..z_max."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    lo, hi = p - r, p + r
    return ((bounds["x_min"] <= lo[:, 0]) & (hi[:, 0] <= bounds["x_max"]) &
            (bounds["y_min"] <= lo[:, 1]) & (hi[:, 1] <= bounds["y_max"]) &
            (bounds["z_min"] <= lo[:, 2]) & (hi[:, 2] <= bounds["z_max"]))


def collides(points, vertices, r):
    """:220-238 for an (n,3) array of query points: any vertex with lo <= v <= hi on all three axes."""
    p = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(len(p), dtype=bool)
    if len(v) == 0:
        return out
    for k in range(len(p)):
        lo, hi = p[k] - r, p[k] + r
        m = (v[:, 2] >= lo[2]) & (v[:, 2] <= hi[2])
        m &= (v[:, 0] >= lo[0]) & (v[:, 0] <= hi[0])
        m &= (v[:, 1] >= lo[1]) & (v[:, 1] <= hi[1])
        out[k] = bool(m.any())
    return out


def grid_axes(bounds, resolution):
    """:123-124 / :177-178."""
    return (np.arange(bounds["x_min"], bounds["x_max"], resolution), np.arange(bounds["y_min"], bounds["y_max"], resolution))


def classify_grid(xs, ys, vertices, bounds, r):
    """State of every grid point (x-major): 0 = robot cube leaves the room (skipped, :136-137), 1 = obstacle, 2 = free."""
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    pts = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, ROBOT_HEIGHT)], axis=1)
    state = np.zeros(len(pts), dtype=np.uint8)
    inb = in_room_bounds(pts, bounds, r)
    # only vertices in the z slab can collide: prefilter once (same comparisons as `collides`)
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
    slab = v[(v[:, 2] >= ROBOT_HEIGHT - r) & (v[:, 2] <= ROBOT_HEIGHT + r)]
    hit = np.zeros(len(pts), dtype=bool)
    idx = np.nonzero(inb)[0]
    hit[idx] = collides(pts[idx], slab, r)
    state[inb & hit] = 1
    state[inb & ~hit] = 2
    return pts, state


def analyze_room_layout(vertices, bounds, r=0.3):
    """:97-202 -> dict(resolution, xs, ys, state, free (n,3), obstacles (m,3), detailed)."""
    dims = np.array([bounds["x_max"] - bounds["x_min"], bounds["y_max"] - bounds["y_min"], bounds["z_max"] - bounds["z_min"]])
    res = max(0.2, min(dims) / 20)
    xs, ys = grid_axes(bounds, res)
    pts, state = classify_grid(xs, ys, vertices, bounds, r)
    detailed = False
    if int((state == 2).sum()) < 10:                         # :151-152
        detailed = True
        res = max(0.15, min(dims) / 30)
        xs, ys = grid_axes(bounds, res)
        pts, state = classify_grid(xs, ys, vertices, bounds, r)
    return dict(resolution=res, xs=xs, ys=ys, state=state, free=pts[state == 2], obstacles=pts[state == 1], detailed=detailed,
                dimensions=dims)


def connectivity_graph(free, r=0.3):
    """:245-258 -> list of ascending neighbour lists."""
    free = np.asarray(free, dtype=np.float64).reshape(-1, 3)
    out = []
    for i in range(len(free)):
        d = free[i] - free
        dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        nb = np.nonzero(dist <= r * 2)[0]
        out.append([int(j) for j in nb if j != i])
    return out


def a_star(start, end, free, graph):
    """:413-473, literally: set-based open list, min() by f-score."""
    if start == end:
        return [start]
    open_set, closed = {start}, set()
    g = {start: 0.0}
    f = {}
    came = {}

    def h(a, b):
        return np.linalg.norm(free[a] - free[b])
    f[start] = h(start, end)
    while open_set:
        cur = min(open_set, key=lambda x: f.get(x, float("inf")))
        if cur == end:
            path = []
            while cur is not None:
                path.append(cur)
                cur = came.get(cur)
            return path[::-1]
        open_set.remove(cur)
        closed.add(cur)
        for nb in graph[cur]:
            if nb in closed:
                continue
            tg = g[cur] + h(cur, nb)
            if nb not in open_set:
                open_set.add(nb)
            elif tg >= g.get(nb, float("inf")):
                continue
            came[nb] = cur
            g[nb] = tg
            f[nb] = tg + h(nb, end)
    return None


def path_cost(path, free):
    return float(sum(np.linalg.norm(free[path[k + 1]] - free[path[k]]) for k in range(len(path) - 1)))


def linear_waypoints(a, b, n):
    """:386-398 -> (n,3)."""
    out = np.zeros((n, 3))
    for i in range(n):
        t = i / (n - 1) if n > 1 else 0
        out[i] = [a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]), a[2] + t * (b[2] - a[2])]
    return out


def waypoints_along_path(path_points, n):
    """:475-527 -> (m,3), m <= n."""
    pp = [np.asarray(p, dtype=np.float64) for p in path_points]
    if len(pp) < 2:
        return np.zeros((0, 3))
    seg = [np.linalg.norm(pp[i + 1] - pp[i]) for i in range(len(pp) - 1)]
    total = 0.0
    for s in seg:
        total += s
    if total < 1e-6:
        return np.array([pp[0]])
    out = []
    for i in range(n):
        if i == n - 1:
            out.append(pp[-1])
            break
        target = (i / (n - 1)) * total
        start = 0.0
        for k, s in enumerate(seg):
            end = start + s
            if target <= end:
                prog = (target - start) / s if s > 0 else 0
                out.append(pp[k] + prog * (pp[k + 1] - pp[k]))
                break
            start = end
    return np.array(out)


def smooth(w, alpha=0.5):
    """:529-554 on an (n,3) array."""
    w = np.asarray(w, dtype=np.float64)
    if len(w) < 3:
        return w.copy()
    out = w.copy()
    for i in range(1, len(w) - 1):
        out[i] = alpha * w[i] + (1 - alpha) * (w[i - 1] + w[i + 1]) / 2
    return out


def count_turns(w):
    """:556-593."""
    if len(w) < 3:
        return 0
    n = 0
    for i in range(1, len(w) - 1):
        v1, v2 = w[i, :2] - w[i - 1, :2], w[i + 1, :2] - w[i, :2]
        n1, n2 = np.linalg.norm(v1), np.linalg.norm(v2)
        if n1 > 1e-6 and n2 > 1e-6:
            ang = np.arccos(np.clip(np.dot(v1 / n1, v2 / n2), -1.0, 1.0))
            if ang > np.pi / 6:
                n += 1
    return n


def path_length(w):
    """:595-612."""
    total = 0.0
    for i in range(1, len(w)):
        total += np.sqrt((w[i, 0] - w[i - 1, 0]) ** 2 + (w[i, 1] - w[i - 1, 1]) ** 2 + (w[i, 2] - w[i - 1, 2]) ** 2)
    return total


def candidate_score(length, smoothness, collisions, min_len):
    """:652-657."""
    return min(length / min_len, 2.0) * 0.4 + smoothness * 0.4 - collisions * 0.1
