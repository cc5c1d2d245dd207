"""CPU: furniture records and collision checks against the fixture the live reference produced
(tests/golden/make_golden_furniture.py; reference trajectory/collision_detector.py, auto_trajectory_generator.py:693-704)."""
import types

import numpy as np


def _detector(lrc_traj, g):
    det = lrc_traj.CollisionDetector(robot_radius=0.35)
    for k in range(5):
        det.add_furniture(lrc_traj.FurnitureInfo(name=f"f{k}", position=g["pos"][k], size=g["size"][k], category="chair"))
    det.add_furniture_from_mesh(types.SimpleNamespace(vertices=g["verts"]), "from_mesh", "table")
    return det


def test_furniture_records_match_the_reference(golden):
    from lrc_b200 import trajectory as T
    g = golden("furniture.npz")
    det = _detector(T, g)
    fl = det.get_furniture_list()
    assert len(fl) == 6 and fl[-1].category == "table"
    assert np.array_equal(fl[-1].position, g["mesh_position"]) and np.array_equal(fl[-1].size, g["mesh_size"])
    keys = ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")
    assert np.array_equal(np.array([[f.get_bounds()[k] for k in keys] for f in fl]), g["bounds"])
    assert np.array_equal(np.array([[f.is_point_inside(p) for f in fl] for p in g["q"]]), g["inside"])
    assert np.array_equal(det._hits(g["q"]), g["bbox"])
    assert g["bbox"].any() and not g["bbox"].all()
    # per-waypoint verdict = first colliding piece in list order (where the reference's loop survives to report one)
    for p, row in zip(g["q"][:60], g["bbox"][:60]):
        hit, f = det.detect_collision(T.Waypoint(x=p[0], y=p[1], z=p[2], yaw=0.0))
        assert hit == bool(row.any()) and (f is None or f.name == fl[int(np.argmax(row))].name)


def test_avoidance_and_statistics_match_the_reference(golden):
    from lrc_b200 import trajectory as T
    g = golden("furniture.npz")
    det = _detector(T, g)
    fl = det.get_furniture_list()
    for k in range(6):
        w = T.Waypoint(x=g["wq"][k, 0], y=g["wq"][k, 1], z=g["wq"][k, 2], yaw=g["wq"][k, 3])
        got = np.array([[a.x, a.y, a.z, a.yaw] for a in det.suggest_avoidance_path(w, fl[k % 6])])
        np.testing.assert_allclose(got, g["avoid"][k], rtol=0, atol=1e-12)
    path = [T.Waypoint(x=p[0], y=p[1], z=p[2], yaw=0.0) for p in g["path"]]
    st = det.get_collision_statistics(path)
    assert st["total_collisions"] == int(g["stats_total"]) == 7 and st["collision_rate"] == float(g["stats_rate"])
    assert st["collision_furniture"] == {"f0": int(g["stats_first"])}
    assert det.get_collision_statistics([]) == {"total_collisions": 0, "collision_rate": 0, "collision_furniture": {}}


def test_planner_furniture_api_is_the_reference_surface():
    """add_furniture / add_furniture_from_mesh / clear_furniture only manage the list (the reference never queries it while
    planning); constructing the generator must not need a GPU."""
    from lrc_b200 import trajectory as T
    gen = T.AutoTrajectoryGenerator(robot_radius=0.25)
    assert gen.collision_detector.robot_radius == 0.25
    gen.add_furniture(T.FurnitureInfo("a", np.zeros(3), np.ones(3), "chair"))
    gen.add_furniture_from_mesh(types.SimpleNamespace(vertices=np.array([[0, 0, 0], [1, 2, 3.0]])), "m")
    gen.add_furniture_from_mesh(types.SimpleNamespace(vertices=np.zeros((0, 3))), "empty")
    names = [f.name for f in gen.collision_detector.get_furniture_list()]
    assert names == ["a", "m"]
    gen.clear_furniture()
    assert gen.collision_detector.get_furniture_list() == []
