"""CPU: the planner oracle (oracle/plan_oracle.py) against fixtures captured from the reference's own
AutoTrajectoryGenerator (tests/golden/make_golden_plan.py)."""
import numpy as np
import pytest

B_KEYS = ("x_min", "x_max", "y_min", "y_max", "z_min", "z_max")


def _scene(lrc, name):
    mesh = lrc.synthetic.box_room(target_tris=6000, seed=3) if name == "plan_room.npz" else lrc.synthetic.planner_tight_room()
    return mesh


@pytest.mark.parametrize("name,detailed", [("plan_room.npz", False), ("plan_tight.npz", True)])
def test_layout_and_graph_match_reference(lrc, golden, name, detailed):
    from oracle import plan_oracle as po
    g = golden(name)
    mesh = _scene(lrc, name)
    b = dict(zip(B_KEYS, g["bounds"]))
    v = mesh.vertices
    assert np.array_equal(g["bounds"], [v[:, 0].min(), v[:, 0].max(), v[:, 1].min(), v[:, 1].max(), v[:, 2].min(), v[:, 2].max()])
    lay = po.analyze_room_layout(v, b)
    assert lay["detailed"] == detailed
    assert np.array_equal(lay["free"], g["free"]) and np.array_equal(lay["obstacles"], g["obstacles"])     # bit-exact, same order
    graph = po.connectivity_graph(lay["free"])
    ptr = np.concatenate([[0], np.cumsum([len(r) for r in graph])])
    assert np.array_equal(ptr, g["graph_ptr"])
    assert np.array_equal(np.array([j for r in graph for j in r], np.int32), g["graph_col"])
    assert float(g["min_trajectory_length"]) == max(lay["dimensions"][0], lay["dimensions"][1]) * 0.2


def test_queries_astar_and_helpers_match_reference(lrc, golden):
    from oracle import plan_oracle as po
    g = golden("plan_room.npz")
    mesh = _scene(lrc, "plan_room.npz")
    b = dict(zip(B_KEYS, g["bounds"]))
    q = g["query/points"]
    assert np.array_equal(po.in_room_bounds(q, b, 0.3), g["query/in_bounds"])
    assert np.array_equal(po.collides(q, mesh.vertices, 0.3), g["query/collides"])
    free = g["free"]
    ptr, col = g["graph_ptr"], g["graph_col"]
    graph = [list(col[ptr[i]:ptr[i + 1]]) for i in range(len(free))]
    for (a, c), cost, ln in list(zip(g["pairs"], g["astar_cost"], g["astar_len"]))[:4]:
        path = po.a_star(int(a), int(c), free, graph)
        assert len(path) == ln and po.path_cost(path, free) == pytest.approx(cost, rel=1e-12)
    pp = [free[i] for i in g["helper/path"]]
    w = po.waypoints_along_path(pp, 40)
    assert np.array_equal(w, g["helper/along"])
    ws = po.smooth(w)
    assert np.array_equal(ws, g["helper/smooth"])
    assert po.count_turns(ws) == int(g["helper/turns"]) and po.count_turns(w) == int(g["helper/turns_raw"])
    assert po.path_length(ws) == float(g["helper/length"])
