"""The oracle's intersector: analytic known answers, float64 cross-check, tie rule and edge cases.
The third-party intersector the reference calls (Open3D/Embree, raycast_engine_cpu.py:46-53) is absent, so these
tests are what pins the restatement."""
import numpy as np
import pytest

MISS = 0xFFFFFFFF


def _rays(o, d):
    o = np.broadcast_to(np.asarray(o, np.float32), np.shape(d))
    return np.concatenate([o, np.asarray(d, np.float32)], axis=1)


def test_single_triangle_known_answer(orc):
    V = np.array([[0, 0, 5], [4, 0, 5], [0, 4, 5]], float)
    F = np.array([[0, 1, 2]], np.int32)
    rays = _rays([[1, 1, 0]], [[0, 0, 1], [0, 0, -1], [0, 0, 2], [3, 3, 5], [1, 0, 0]])
    for fn in (lambda r: orc.OracleScene((V, F)).cast_rays(r), lambda r: orc.cast_rays_brute((V, F), r)):
        t, pid = fn(rays)
        assert t[0] == 5.0 and pid[0] == 0            # straight up
        assert t[1] == np.inf and pid[1] == MISS      # pointing away: t < 0 is not a hit
        assert t[2] == 2.5 and pid[2] == 0            # t is in units of |d| (direction used as given)
        assert pid[3] == MISS                         # lands at (4,4,5): outside u+v<=1
        assert pid[4] == MISS                         # parallel to the plane: det == 0


def test_two_sided_and_closest(orc):
    V = np.array([[0, 0, 2], [1, 0, 2], [0, 1, 2], [0, 0, 1], [0, 1, 1], [1, 0, 1]], float)   # second one is wound the other way
    F = np.array([[0, 1, 2], [3, 4, 5]], np.int32)
    t, pid = orc.OracleScene((V, F)).cast_rays(_rays([[0.2, 0.2, 0]], [[0, 0, 1]]))
    assert t[0] == 1.0 and pid[0] == 1
    t, pid = orc.OracleScene((V, F)).cast_rays(_rays([[0.2, 0.2, 3]], [[0, 0, -1]]))
    assert t[0] == 1.0 and pid[0] == 0


def test_equal_t_tie_goes_to_smallest_triangle_id(orc):
    tri = [[0, 0, 1], [1, 0, 1], [0, 1, 1]]
    V = np.array(tri * 3, float)
    F = np.arange(9, dtype=np.int32).reshape(3, 3)[::-1].copy()      # ids 0,1,2 all coincide geometrically
    for fn in (lambda r: orc.OracleScene((V, F)).cast_rays(r), lambda r: orc.cast_rays_brute((V, F), r)):
        t, pid = fn(_rays([[0.25, 0.25, 0]], [[0, 0, 1]]))
        assert t[0] == 1.0 and pid[0] == 0


def test_origin_on_surface_counts_as_hit_at_zero(orc):
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    F = np.array([[0, 1, 2]], np.int32)
    t, pid = orc.OracleScene((V, F)).cast_rays(_rays([[0.25, 0.25, 0]], [[0, 0, 1]]))
    assert t[0] == 0.0 and pid[0] == 0                                # Embree contract: tnear = 0 is inclusive


def test_empty_scene_and_no_rays(orc):
    sc = orc.OracleScene((np.zeros((0, 3)), np.zeros((0, 3), np.int32)))
    t, pid = sc.cast_rays(_rays([[0, 0, 0]], [[1, 0, 0]]))
    assert t[0] == np.inf and pid[0] == MISS
    V = np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], float)
    t, pid = orc.OracleScene((V, np.array([[0, 1, 2]], np.int32))).cast_rays(np.zeros((0, 6), np.float32))
    assert t.shape == (0,) and pid.shape == (0,)


def test_degenerate_triangles_never_hit(orc):
    V = np.array([[0, 0, 1], [1, 1, 1], [2, 2, 1], [0, 0, 2], [0, 0, 2], [0, 0, 2]], float)   # collinear, point
    F = np.array([[0, 1, 2], [3, 4, 5]], np.int32)
    t, pid = orc.OracleScene((V, F)).cast_rays(_rays([[0.5, 0.5, 0]], [[0, 0, 1], [-0.5, -0.5, 2]]))
    assert np.all(pid == MISS) and np.all(np.isinf(t))


def test_analytic_empty_box(orc, lrc):
    """Sensor inside an axis-aligned box: t = min over the six planes of (plane - o)/d, in closed form."""
    lo, hi = np.array([0, 0, 0.0]), np.array([4, 3, 2.5])
    mesh = lrc.synthetic.empty_box(tuple(lo), tuple(hi), pitch=0.5)
    pose = np.eye(4)
    pose[:3, 3] = (1.3137, 0.9271, 1.1077)                     # generic position: no ray runs through a mesh edge
    rays = orc.gen_rays_single_axis(pose, [31.0, 10.3, 0.0, -10.7, -44.0], 720)
    t, pid = orc.OracleScene(mesh).cast_rays(rays)
    assert (pid != MISS).all()
    o, d = rays[:, :3].astype(np.float64), rays[:, 3:].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        tt = np.where(d > 0, (hi - o) / d, np.where(d < 0, (lo - o) / d, np.inf))
    np.testing.assert_allclose(t, tt.min(1), rtol=0, atol=2e-6)
    # the winning plane must own the triangle: its label says floor / ceiling / wall
    sem = mesh.triangle_labels[pid] & 0xFFFF
    axis = tt.argmin(1)
    expect = np.where(axis == 2, np.where(d[:, 2] < 0, 1, 0), 2)
    assert np.array_equal(sem, expect)


def test_documented_edge_leak(orc, lrc):
    """A ray aimed EXACTLY at a shared mesh edge can slip between the two float32 Moller-Trumbore tests (both
    reject by one rounding).  This is the 'grazing-edge tie' class of BASELINE.json (budget 1e-5 of rays): the
    float64 evaluation hits, the float32 spec may miss -- and the CUDA engine must reproduce the spec bit for bit."""
    mesh = lrc.synthetic.empty_box((0, 0, 0), (4, 3, 2.5), pitch=0.5)
    ray = _rays([[1.3, 0.9, 1.1]], [[4.3297803e-17, 0.70710677, -0.70710677]])    # lands on the edge y = 2.0 of the floor grid
    t64, p64 = orc.cast_rays_brute_f64(mesh, ray)
    assert p64[0] != MISS and t64[0] == pytest.approx(1.1 * np.sqrt(2), abs=1e-6)
    t32, p32 = orc.cast_rays_brute(mesh, ray)
    tb, pb = orc.OracleScene(mesh).cast_rays(ray)
    assert p32[0] == pb[0] and (t32[0] == tb[0])                                   # BVH == exhaustive, whatever the verdict


def test_bvh_equals_exhaustive_and_float64_on_box_room(orc, lrc):
    """C1: 16000 rays x 50k triangles.  BVH result == exhaustive float32 result bit for bit; float64 evaluation
    of the same formulas agrees on every triangle id (ties/edges budget 1e-5) and to 1e-4 m in distance."""
    mesh = lrc.synthetic.box_room()
    rays = orc.gen_rays_single_axis(lrc.synthetic.box_room_pose(), lrc.Indoor8LineLidarIntrinsics().vertical_degrees, 2000)
    t, pid = orc.OracleScene(mesh).cast_rays(rays)
    tb, pb = orc.cast_rays_brute(mesh, rays)
    assert np.array_equal(pid, pb) and np.array_equal(t, tb)
    t64, p64 = orc.cast_rays_brute_f64(mesh, rays)
    assert (pid != p64).mean() <= 1e-5
    both = (pid != MISS) & (p64 != MISS)
    assert both.mean() > 0.999
    assert np.abs(t[both] - t64[both]).max() <= 1e-4


def test_unnormalised_directions_scale_t(orc, lrc):
    mesh = lrc.synthetic.box_room(target_tris=3000)
    rays = orc.gen_rays_single_axis(lrc.synthetic.box_room_pose(), [5.0, -12.0], 64)
    t1, p1 = orc.OracleScene(mesh).cast_rays(rays)
    scaled = rays.copy()
    scaled[:, 3:] *= 4.0                                                           # exact in float32
    t4, p4 = orc.OracleScene(mesh).cast_rays(scaled)
    assert np.array_equal(p1, p4)
    np.testing.assert_allclose(t4 * 4.0, t1, rtol=1e-6)
