"""
Golden fixture for the furniture records / collision checks, produced by the LIVE REFERENCE's own
trajectory/collision_detector.py (imported with a stub ``open3d``).  Run once in the build container:

    python tests/golden/make_golden_furniture.py      ->  tests/golden/furniture.npz

Only the parts of the reference that can run are captured: bounds, point-in-box, the expanded-box test
(``_check_bbox_collision``), ``add_furniture_from_mesh``, ``suggest_avoidance_path`` and the path statistics of a path on
which every waypoint collides with the FIRST piece (the reference's ``detect_collision`` raises AttributeError as soon as
a box test fails, because ``FurnitureInfo`` has no ``mesh`` attribute, collision_detector.py:125).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
o3d = types.ModuleType("open3d")
o3d.geometry = types.SimpleNamespace(TriangleMesh=object, PointCloud=object, AxisAlignedBoundingBox=object)
o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: a)
sys.modules["open3d"] = o3d
sys.path.insert(0, "/root/reference")
from trajectory.collision_detector import CollisionDetector, FurnitureInfo  # noqa: E402
from trajectory.trajectory_generator import Waypoint  # noqa: E402


def main():
    rng = np.random.default_rng(42)
    det = CollisionDetector(robot_radius=0.35)
    pos = rng.uniform(0.5, 6.0, (5, 3))
    size = rng.uniform(0.3, 1.5, (5, 3))
    for k in range(5):
        det.add_furniture(FurnitureInfo(name=f"f{k}", position=pos[k], size=size[k], category="chair"))
    verts = rng.uniform(1.0, 3.0, (200, 3))
    det.add_furniture_from_mesh(types.SimpleNamespace(vertices=verts), "from_mesh", "table")
    fl = det.get_furniture_list()
    bounds = np.array([[f.get_bounds()[k] for k in ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")] for f in fl])
    q = rng.uniform(0.0, 7.0, (400, 3))
    inside = np.array([[f.is_point_inside(p) for f in fl] for p in q])
    bbox = np.array([[det._check_bbox_collision(p, f) for f in fl] for p in q])
    # avoidance suggestions for a few (waypoint, furniture) pairs
    wq = rng.uniform(0.5, 6.0, (6, 4))
    avoid = []
    for k in range(6):
        ws = det.suggest_avoidance_path(Waypoint(x=wq[k, 0], y=wq[k, 1], z=wq[k, 2], yaw=wq[k, 3]), fl[k % len(fl)])
        avoid.append([[w.x, w.y, w.z, w.yaw] for w in ws])
    # a path entirely inside the first piece's expanded box: the one case the reference's detect_path_collision survives
    f0 = fl[0]
    path = [Waypoint(x=f0.position[0] + 0.01 * i, y=f0.position[1], z=f0.position[2], yaw=0.0) for i in range(7)]
    stats = det.get_collision_statistics(path)
    np.savez(os.path.join(HERE, "furniture.npz"), pos=pos, size=size, verts=verts, mesh_position=fl[-1].position, mesh_size=fl[-1].size,
             bounds=bounds, q=q, inside=inside, bbox=bbox, wq=wq, avoid=np.array(avoid),
             path=np.array([[w.x, w.y, w.z] for w in path]), stats_total=stats["total_collisions"], stats_rate=stats["collision_rate"],
             stats_first=stats["collision_furniture"]["f0"])
    print("furniture.npz written:", bounds.shape, inside.sum(), bbox.sum(), stats)


if __name__ == "__main__":
    main()
