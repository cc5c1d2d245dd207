"""CPU: the host-side halves of the coverage planner -- the native A* (lrc_astar is host code, no GPU needed) and the
polyline helpers -- against fixtures captured from the reference's AutoTrajectoryGenerator."""
import ctypes as C

import numpy as np
import pytest


def _astar(free, ptr, col, a, b):
    from lrc_b200 import _native as nat
    lib = nat.load()
    free = np.ascontiguousarray(free, np.float64)
    ptr, col = np.ascontiguousarray(ptr, np.int32), np.ascontiguousarray(col, np.int32)
    n = len(free)
    path = np.empty(n, np.int32)
    ln, cost = C.c_int32(0), C.c_double(0)
    rc = lib.lrc_astar(C.c_void_p(ptr.ctypes.data), C.c_void_p(col.ctypes.data), C.c_void_p(free.ctypes.data), n, int(a), int(b),
                       C.c_void_p(path.ctypes.data), n, C.byref(ln), C.byref(cost))
    assert rc == 0
    return [int(i) for i in path[:ln.value]], cost.value


@pytest.mark.parametrize("name", ["plan_room.npz", "plan_tight.npz"])
def test_native_astar_cost_equals_reference(golden, name):
    g = golden(name)
    free, ptr, col = g["free"], g["graph_ptr"], g["graph_col"]
    same_shape = 0
    for (a, b), want_cost, want_len in zip(g["pairs"], g["astar_cost"], g["astar_len"]):
        path, cost = _astar(free, ptr, col, a, b)
        if want_len == 0:
            assert path == []
            continue
        assert path[0] == a and path[-1] == b
        for u, v in zip(path[:-1], path[1:]):                         # every hop is an edge of the reference's graph
            assert v in col[ptr[u]:ptr[u + 1]]
        walked = sum(np.linalg.norm(free[v] - free[u]) for u, v in zip(path[:-1], path[1:]))
        assert cost == pytest.approx(want_cost, rel=1e-9) and walked == pytest.approx(want_cost, rel=1e-9)
        same_shape += len(path) == want_len
    assert same_shape >= 1


def test_native_astar_edge_cases():
    free = np.array([[0, 0, 1], [0.5, 0, 1], [5, 5, 1]], float)
    ptr, col = np.array([0, 1, 2, 2], np.int32), np.array([1, 0], np.int32)
    assert _astar(free, ptr, col, 0, 0) == ([0], 0.0)
    assert _astar(free, ptr, col, 0, 1) == ([0, 1], 0.5)
    assert _astar(free, ptr, col, 0, 2)[0] == []                      # disconnected: no path (reference returns None)


def test_polyline_helpers_match_reference(lrc, golden):
    g = golden("plan_room.npz")
    A = lrc.trajectory.AutoTrajectoryGenerator
    pp = [g["free"][i] for i in g["helper/path"]]
    w = A._generate_waypoints_along_path(pp, 40)
    assert np.array_equal(w, g["helper/along"])
    ws = A._smooth_trajectory(w)
    assert np.array_equal(ws, g["helper/smooth"])
    assert A._count_turns(ws) == int(g["helper/turns"]) and A._count_turns(w) == int(g["helper/turns_raw"])
    assert A._calculate_trajectory_length(ws) == float(g["helper/length"])
    lin = A._generate_linear_waypoints(pp[0], pp[-1], 5)
    assert np.array_equal(lin[0], pp[0]) and np.allclose(lin[-1], pp[-1], rtol=0, atol=1e-12) and lin.shape == (5, 3)
    wps = [lrc.Waypoint(p[0], p[1], p[2], 0) for p in ws]
    assert A._calculate_smoothness_score(wps) == 1.0
