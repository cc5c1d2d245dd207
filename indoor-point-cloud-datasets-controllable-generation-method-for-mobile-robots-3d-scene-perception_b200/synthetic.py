"""
Procedural, seeded stand-ins for the NKSR room meshes the reference ray-casts (S3DIS data and the NKSR
reconstructor are unavailable): the workloads of BASELINE.json's configs, SURVEY.md section 8d.

    box_room()      C1  10 x 8 x 3 m room + 6 furniture boxes, ~50k triangles
    office()        C2/C3  20 x 15 x 3 m office with ~60 desks / chairs / shelves, ~1M triangles (~4 cm facets)
    floor_plan()    C4  60 x 40 x 3 m floor of 24 rooms with door openings and furniture, ~5M triangles

Every surface is a uniform grid of quads split into two triangles.  Interior grid vertices are jittered by
N(0, 1 mm) so that exact ties between coplanar facets have measure zero, while vertices on the rim of a face
stay put so that adjacent faces meet without cracks.  Per-triangle labels use the S3DIS class ids
(reference s3dis_annotation_loader.py:51-65): 0 ceiling, 1 floor, 2 wall, 7 table, 8 chair, 10 bookcase.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

from .core import TriangleMesh, pack_labels
from .trajectory import Waypoint, polyline_waypoints

SEM_CEILING, SEM_FLOOR, SEM_WALL, SEM_TABLE, SEM_CHAIR, SEM_BOOKCASE = 0, 1, 2, 7, 8, 10


@dataclass
class Box:
    lo: Tuple[float, float, float]
    hi: Tuple[float, float, float]
    sem: int                 # semantic id of the side faces (and of all faces for furniture)
    ins: int                 # instance id
    room_shell: bool = False  # floor / ceiling / wall semantics per face


def _face_counts(box: Box, pitch: float) -> np.ndarray:
    ext = np.asarray(box.hi, float) - np.asarray(box.lo, float)
    return np.maximum(1, np.ceil(ext / pitch - 1e-9).astype(int))


def _box_triangle_count(box: Box, pitch: float) -> int:
    n = _face_counts(box, pitch)
    return int(4 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2]))


def _tessellate(boxes: Sequence[Box], pitch: float, jitter: float, rng: np.random.Generator):
    verts, tris, sems, inss = [], [], [], []
    base = 0
    for b in boxes:
        lo, hi = np.asarray(b.lo, float), np.asarray(b.hi, float)
        n = _face_counts(b, pitch)
        axes = [np.linspace(lo[a], hi[a], n[a] + 1) for a in range(3)]
        for axis in range(3):
            u, v = [a for a in range(3) if a != axis]
            for side in (0, 1):
                gu, gv = np.meshgrid(axes[u], axes[v], indexing="ij")
                p = np.empty(gu.shape + (3,))
                p[..., u], p[..., v] = gu, gv
                p[..., axis] = lo[axis] if side == 0 else hi[axis]
                if jitter > 0 and n[u] > 1 and n[v] > 1:
                    p[1:-1, 1:-1] += rng.normal(0.0, jitter, size=p[1:-1, 1:-1].shape)
                nu, nv = n[u], n[v]
                iu, iv = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
                a00 = (iu * (nv + 1) + iv).ravel() + base
                a01 = a00 + 1
                a10 = a00 + (nv + 1)
                a11 = a10 + 1
                t = np.empty((2 * nu * nv, 3), np.int64)
                t[0::2] = np.stack([a00, a10, a11], 1)
                t[1::2] = np.stack([a00, a11, a01], 1)
                verts.append(p.reshape(-1, 3))
                tris.append(t)
                if b.room_shell:
                    sem = SEM_WALL if axis != 2 else (SEM_FLOOR if side == 0 else SEM_CEILING)
                    ins = b.ins + (0 if axis != 2 else (1 if side == 0 else 2))
                else:
                    sem, ins = b.sem, b.ins
                sems.append(np.full(len(t), sem, np.uint32))
                inss.append(np.full(len(t), ins, np.uint32))
                base += (nu + 1) * (nv + 1)
    V = np.concatenate(verts).astype(np.float32).astype(np.float64)   # what a float32 PLY would hold
    F = np.concatenate(tris).astype(np.int32)
    return V, F, pack_labels(np.concatenate(sems), np.concatenate(inss))


def _fit_pitch(boxes: Sequence[Box], target: int) -> float:
    lo, hi = 1e-3, 5.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if sum(_box_triangle_count(b, mid) for b in boxes) > target:
            lo = mid
        else:
            hi = mid
    return hi


def build_mesh(boxes: Sequence[Box], target_tris: int, seed: int = 0, jitter: float = 1e-3) -> TriangleMesh:
    pitch = _fit_pitch(boxes, target_tris)
    V, F, lab = _tessellate(boxes, pitch, jitter, np.random.default_rng(seed))
    mesh = TriangleMesh(V, F, lab)
    mesh.pitch = pitch
    return mesh


# ---- C1 ------------------------------------------------------------------------------------------
def box_room(target_tris: int = 50_000, seed: int = 0) -> TriangleMesh:
    boxes = [Box((0, 0, 0), (10, 8, 3), SEM_WALL, 0, room_shell=True)]
    furn = [((1.0, 1.0, 0.0), (2.6, 1.8, 0.75), SEM_TABLE), ((6.0, 1.2, 0.0), (7.6, 2.0, 0.75), SEM_TABLE),
            ((4.6, 5.0, 0.0), (5.1, 5.5, 0.9), SEM_CHAIR), ((7.5, 5.5, 0.0), (8.0, 6.0, 0.9), SEM_CHAIR),
            ((0.1, 4.0, 0.0), (0.5, 6.0, 2.0), SEM_BOOKCASE), ((4.0, 7.5, 0.0), (6.0, 7.9, 2.0), SEM_BOOKCASE)]
    for k, (lo, hi, sem) in enumerate(furn):
        boxes.append(Box(lo, hi, sem, 3 + k))
    return build_mesh(boxes, target_tris, seed)


def planner_tight_room(seed: int = 5) -> TriangleMesh:
    """A 1.4 x 1.6 x 2.5 m closet with a pillar: the planner's coarse 0.2 m grid finds fewer than 10 free points, which
    sends the reference into its detailed-resolution branch (auto_trajectory_generator.py:151-152,167-202)."""
    boxes = [Box((0, 0, 0), (1.4, 1.6, 2.5), SEM_WALL, 0, room_shell=True),
             Box((0.55, 0.75, 0.0), (0.85, 0.85, 2.0), SEM_BOOKCASE, 3)]
    return build_mesh(boxes, 3000, seed)


def box_room_pose() -> np.ndarray:
    """C1's single pose: (3.137, 2.718, 1.0), yaw 0.3 rad."""
    return Waypoint(3.137, 2.718, 1.0, 0.3).to_pose_matrix()


# ---- C2 / C3 -------------------------------------------------------------------------------------
_OFFICE_ROWS = (2.5, 5.0, 7.5, 10.0, 12.5)          # desk rows (y); aisles run between them
_OFFICE_AISLES = (1.25, 3.75, 6.25, 8.75, 11.25, 13.75)


def office_boxes(seed: int = 0) -> List[Box]:
    rng = np.random.default_rng(seed + 1000)
    boxes = [Box((0, 0, 0), (20, 15, 3), SEM_WALL, 0, room_shell=True)]
    ins = 3
    for y in _OFFICE_ROWS:
        for x in np.arange(2.0, 18.5, 2.75):          # 6 desk + chair pairs per row
            w, d = 1.4 + 0.4 * rng.random(), 0.7 + 0.2 * rng.random()
            boxes.append(Box((x, y - d / 2, 0.0), (x + w, y + d / 2, 0.72 + 0.06 * rng.random()), SEM_TABLE, ins)); ins += 1
            cx = x + w / 2 + 0.2 * (rng.random() - 0.5)
            boxes.append(Box((cx - 0.25, y + d / 2 + 0.05, 0.0), (cx + 0.25, y + d / 2 + 0.5, 0.85 + 0.1 * rng.random()), SEM_CHAIR, ins)); ins += 1
    # bookcases against the two short walls, between aisles
    for y in _OFFICE_ROWS:
        boxes.append(Box((0.05, y - 0.9, 0.0), (0.45, y + 0.9, 1.9 + 0.2 * rng.random()), SEM_BOOKCASE, ins)); ins += 1
        boxes.append(Box((19.55, y - 0.9, 0.0), (19.95, y + 0.9, 1.9 + 0.2 * rng.random()), SEM_BOOKCASE, ins)); ins += 1
    return boxes


def office(target_tris: int = 1_000_000, seed: int = 0) -> TriangleMesh:
    return build_mesh(office_boxes(seed), target_tris, seed)


def office_waypoints(count: int = 100) -> List[Waypoint]:
    """Lawnmower through the aisles at z = 1.0, yaw = 0 (reference auto_trajectory_generator.py:122,400)."""
    path = []
    for k, y in enumerate(_OFFICE_AISLES):
        xs = (1.0, 19.0) if k % 2 == 0 else (19.0, 1.0)
        path += [(xs[0], y), (xs[1], y)]
    return polyline_waypoints(path, count, z=1.0, yaw=0.0)


# ---- C4 ------------------------------------------------------------------------------------------
def floor_plan_boxes(seed: int = 0, rooms_x: int = 6, rooms_y: int = 4, room: float = 10.0, wall: float = 0.2,
                     door: float = 1.2) -> List[Box]:
    rng = np.random.default_rng(seed + 2000)
    X, Y, Hh = rooms_x * room, rooms_y * room, 3.0
    boxes = [Box((0, 0, 0), (X, Y, Hh), SEM_WALL, 0, room_shell=True)]
    ins = 3
    # interior walls parallel to y (between room columns), door centred on each room's y-centre
    for i in range(1, rooms_x):
        x = i * room
        for j in range(rooms_y):
            y0, yc, y1 = j * room, j * room + room / 2, (j + 1) * room
            boxes.append(Box((x - wall / 2, y0, 0), (x + wall / 2, yc - door / 2, Hh), SEM_WALL, ins)); ins += 1
            boxes.append(Box((x - wall / 2, yc + door / 2, 0), (x + wall / 2, y1, Hh), SEM_WALL, ins)); ins += 1
    # interior walls parallel to x (between room rows), door centred on each room's x-centre
    for j in range(1, rooms_y):
        y = j * room
        for i in range(rooms_x):
            x0, xc, x1 = i * room, i * room + room / 2, (i + 1) * room
            boxes.append(Box((x0, y - wall / 2, 0), (xc - door / 2, y + wall / 2, Hh), SEM_WALL, ins)); ins += 1
            boxes.append(Box((xc + door / 2, y - wall / 2, 0), (x1, y + wall / 2, Hh), SEM_WALL, ins)); ins += 1
    # furniture: kept clear of each room's two centre lines (the trajectory runs along them)
    for i in range(rooms_x):
        for j in range(rooms_y):
            ox, oy = i * room, j * room
            for (qx, qy) in ((1.0, 1.0), (6.2, 1.0), (1.0, 6.2), (6.2, 6.2)):
                w, d = 1.4 + 0.8 * rng.random(), 0.7 + 0.5 * rng.random()
                x0, y0 = ox + qx + 0.8 * rng.random(), oy + qy + 0.8 * rng.random()
                kind = rng.integers(0, 3)
                sem, h = ((SEM_TABLE, 0.75), (SEM_CHAIR, 0.9), (SEM_BOOKCASE, 2.0))[kind]
                if sem == SEM_CHAIR:
                    w = d = 0.5
                boxes.append(Box((x0, y0, 0.0), (x0 + w, y0 + d, h), sem, ins)); ins += 1
    return boxes


def floor_plan(target_tris: int = 5_000_000, seed: int = 0) -> TriangleMesh:
    return build_mesh(floor_plan_boxes(seed), target_tris, seed)


def floor_plan_waypoints(count: int = 500, rooms_x: int = 6, rooms_y: int = 4, room: float = 10.0) -> List[Waypoint]:
    """Snake through every row of rooms along the door axis, z = 1.0, yaw = 0."""
    X = rooms_x * room
    path = []
    for j in range(rooms_y):
        y = j * room + room / 2
        xs = (1.0, X - 1.0) if j % 2 == 0 else (X - 1.0, 1.0)
        path += [(xs[0], y), (xs[1], y)]
        if j + 1 < rooms_y:   # move to the next row through the door above the end room's centre
            xc = (room / 2) if xs[1] < X / 2 else X - room / 2
            path += [(xc, y), (xc, y + room)]
    return polyline_waypoints(path, count, z=1.0, yaw=0.0)


# ---- tiny analytic scenes for known-answer tests --------------------------------------------------
def empty_box(lo=(0.0, 0.0, 0.0), hi=(4.0, 3.0, 2.5), pitch: float = 0.5) -> TriangleMesh:
    """Un-jittered axis-aligned box: a ray from inside hits the nearest of six planes in closed form."""
    V, F, lab = _tessellate([Box(lo, hi, SEM_WALL, 0, room_shell=True)], pitch, 0.0, np.random.default_rng(0))
    return TriangleMesh(V, F, lab)
