"""The oracle's frame-level wrapper vs the reference's UNMODIFIED RaycastEngineCPU code run on top of the same
intersector (fixtures: tests/golden/frame_box_room.npz, generator: tests/golden/make_golden.py)."""
import numpy as np
import pytest


class _Lidar:
    """Duck-typed sensor: exactly what the reference engine touches (raycast_engine_cpu.py:91,95,97)."""

    def __init__(self, rays, pose, max_range):
        self._rays, self.pose = rays, pose
        self.intrinsics = type("I", (), {"max_range": max_range})()

    def get_rays(self):
        return self._rays


@pytest.fixture(scope="module")
def frame(golden):
    return golden("frame_box_room.npz")


@pytest.fixture(scope="module")
def mesh(frame):
    return (frame["verts"], frame["tris"])


def test_8line_frame_matches_reference_code(orc, frame, mesh, golden_poses):
    rays = orc.gen_rays_single_axis(golden_poses["posed"], [15, 10, 5, 0, -5, -10, -15, -20], 2000)
    pts, inc = orc.OracleEngineCPU().lidar_intersect_mesh(_Lidar(rays, golden_poses["posed"], 20.0), mesh)
    assert pts.dtype == np.float32 and inc.dtype == np.float64
    assert np.array_equal(pts, frame["8line/points"])
    np.testing.assert_allclose(inc, frame["8line/incident"], rtol=0, atol=1e-12)


def test_range_filter_is_strict_and_in_float64(orc, frame, mesh, golden_poses):
    rays = orc.gen_rays_single_axis(golden_poses["posed"], [15, 10, 5, 0, -5, -10, -15, -20], 500)
    pts, inc = orc.OracleEngineCPU().lidar_intersect_mesh(_Lidar(rays, golden_poses["posed"], 3.0), mesh)
    assert 0 < len(pts) < len(rays)
    assert np.array_equal(pts, frame["short/points"])
    np.testing.assert_allclose(inc, frame["short/incident"], rtol=0, atol=1e-12)
    assert np.all(np.linalg.norm(pts - golden_poses["posed"][:3, 3], axis=1) < 3.0)


def test_misses_are_compacted_in_ray_order(orc, frame, mesh):
    rays = frame["outside/rays"]
    eng = orc.OracleEngineCPU()
    pts = eng.rays_intersect_mesh(rays=rays, mesh=mesh)
    assert 0 < len(pts) < len(rays)
    assert np.array_equal(pts, frame["outside/points"])


def test_c_epilogue_equals_numpy_restatement(orc, frame, mesh, golden_poses):
    """orc_epilogue (C, used for the multi-threaded CPU figure) == the numpy restatement == the reference."""
    pose = golden_poses["posed"]
    rays = orc.gen_rays_single_axis(pose, [15, 10, 5, 0, -5, -10, -15, -20], 2000)
    sc = orc.OracleScene(mesh)
    t, pid = sc.cast_rays(rays)
    fr = orc.epilogue_c(rays, t, pid, center=pose[:3, 3], max_range=20.0, tri_label=frame["labels"])
    assert np.array_equal(fr.points, frame["8line/points"])
    np.testing.assert_allclose(fr.incident, frame["8line/incident"], rtol=0, atol=1e-12)
    assert np.array_equal(fr.label, frame["labels"][fr.prim_id])
    assert np.all(np.diff(fr.ray_idx.astype(np.int64)) > 0)
    eng = orc.OracleEngineCPU()
    eng.lidar_intersect_mesh(_Lidar(rays, pose, 20.0), mesh)
    assert np.array_equal(eng.last_prim_id, fr.prim_id)
    # rays_intersect_mesh flavour: no range filter
    fr2 = orc.epilogue_c(frame["outside/rays"], *sc.cast_rays(frame["outside/rays"]))
    assert np.array_equal(fr2.points, frame["outside/points"])


def test_engine_input_checks(orc, mesh):
    eng = orc.OracleEngineCPU()
    with pytest.raises(TypeError):
        eng.rays_intersect_mesh(rays=[[0, 0, 0, 1, 0, 0]], mesh=mesh)          # reference :40-41
    with pytest.raises(ValueError):
        eng.rays_intersect_mesh(rays=np.zeros((4, 5), np.float32), mesh=mesh)  # reference :42-43


def test_empty_frame(orc, mesh):
    pose = np.eye(4)
    pose[:3, 3] = (100.0, 100.0, 100.0)
    rays = np.concatenate([np.tile(pose[:3, 3], (4, 1)), np.tile([0, 0, 1.0], (4, 1))], 1).astype(np.float32)
    pts, inc = orc.OracleEngineCPU().lidar_intersect_mesh(_Lidar(rays, pose, 20.0), mesh)
    assert pts.shape == (0, 3) and inc.shape == (0,)
