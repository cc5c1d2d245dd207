"""Host-side mirror of the reference interface (no GPU needed): names, fields, defaults, factories, pose format."""
import dataclasses

import numpy as np
import pytest


def test_public_names_match_reference_modules(lrc):
    # reference lidar/__init__.py:6-16 and raycast_engine/__init__.py:6-14
    for name in ("LidarIntrinsics", "Indoor8LineLidarIntrinsics", "DualAxisLidarIntrinsics", "IndoorLidar",
                 "DualAxisLidar", "create_lidar"):
        assert hasattr(lrc.lidar, name)
    for name in ("RaycastEngineBase", "RaycastEngineGPU"):
        assert hasattr(lrc.raycast_engine, name)
    assert issubclass(lrc.RaycastEngineGPU, lrc.RaycastEngineBase)
    for m in ("rays_intersect_mesh", "lidar_intersect_mesh"):
        assert m in lrc.RaycastEngineBase.__abstractmethods__


def test_indoor_8line_defaults(lrc):
    """Values of reference lidar/lidar_intrinsics.py:222-243."""
    i = lrc.Indoor8LineLidarIntrinsics.create_standard_8line()
    expect = dict(fov_up=15.0, fov_down=20.0, vertical_res=8, horizontal_res=2000, max_range=20.0,
                  vertical_degrees=[15, 10, 5, 0, -5, -10, -15, -20], min_range=0.1, range_resolution=0.01,
                  scan_frequency=10.0, points_per_beam=2000, range_noise_std=0.02, angle_noise_std=0.01, dual_axis=False,
                  capture_rate=200000, intensity_noise_std=0.1, dropout_probability=0.05)
    assert dataclasses.asdict(i) == expect
    assert [f.name for f in dataclasses.fields(i)] == list(expect)           # positional order is API too
    assert i.get_total_points_per_scan() == 16000 and i.get_range_limits() == (0.1, 20.0)


def test_indoor_factories(lrc):
    I = lrc.Indoor8LineLidarIntrinsics
    d = I.create_dense_32line()                                                # reference :270-289
    assert (d.vertical_res, d.horizontal_res, d.max_range, d.points_per_beam) == (32, 4000, 25.0, 3000)
    assert d.vertical_degrees[0] == 15.0 and d.vertical_degrees[-1] == -20.0 and len(d.vertical_degrees) == 32
    assert d.vertical_degrees[1] == round(15.0 - 35.0 / 31.0, 1)
    assert (d.range_resolution, d.range_noise_std, d.angle_noise_std) == (0.005, 0.01, 0.005)
    b = I.create_leica_blk2go()                                                # :292-317
    assert (b.vertical_res, b.horizontal_res, b.min_range, b.scan_frequency, b.dual_axis, b.capture_rate) == (64, 8000, 0.5, 20.0, True, 420000)
    assert b.get_total_points_per_scan() == 512000
    h = I.create_high_resolution_8line()                                       # :251-257
    assert (h.horizontal_res, h.points_per_beam, h.range_resolution) == (4000, 4000, 0.005)
    l = I.create_low_cost_8line()                                              # :260-267
    assert (l.horizontal_res, l.points_per_beam, l.range_resolution, l.range_noise_std) == (1000, 1000, 0.02, 0.05)
    c = I.create_custom_lidar(num_beams=3, beam_angles=[22.5, -1.25, -40.0], horizontal_resolution=0.7)   # :320-350
    assert (c.fov_up, c.fov_down, c.vertical_res, c.horizontal_res) == (22.5, 40.0, 3, 514)
    assert I.create_custom_lidar(horizontal_resolution=0.01).horizontal_res == 10000


def test_dual_axis_defaults_and_blk2go(lrc):
    d = lrc.DualAxisLidarIntrinsics()
    assert (d.point_rate, d.scan_duration, d.num_vertical_lines, d.angle_noise_std, d.dropout_probability) == (420000, 1.0, 32, 0.001, 0.02)
    b = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()                  # reference :152-186
    assert (b.point_rate, b.scan_duration, b.frame_duration, b.max_range) == (640000, 0.1, 0.1, 25.0)
    assert b.get_total_points_per_scan() == 64000 and b.get_scan_frequency() == pytest.approx(10.0)
    assert b.theta_range == (-20.0 * np.pi / 180, 15.0 * np.pi / 180) and b.swing_amplitude == 5.0 * np.pi / 180
    assert b.get_range_limits() == (0.5, 25.0)
    assert lrc.rays_per_frame(b) == 64000
    assert lrc.DualAxisLidarIntrinsics.create_custom_dual_axis(point_rate=1000).point_rate == 1000


def test_create_lidar_dispatch_and_errors(lrc):
    pose = np.eye(4)
    assert isinstance(lrc.create_lidar(lrc.Indoor8LineLidarIntrinsics(), pose), lrc.IndoorLidar)
    assert isinstance(lrc.create_lidar(lrc.DualAxisLidarIntrinsics(), pose), lrc.DualAxisLidar)
    with pytest.raises(ValueError):
        lrc.create_lidar(object(), pose)                                       # reference indoor_lidar.py:392-393
    with pytest.raises(AssertionError):
        lrc.IndoorLidar(lrc.Indoor8LineLidarIntrinsics(), np.eye(3))           # reference :21-25
    with pytest.raises(AssertionError):
        lrc.IndoorLidar(lrc.DualAxisLidarIntrinsics(), pose)
    assert lrc.get_lidar_type(lrc.Indoor8LineLidarIntrinsics.create_dense_32line()) == "32-line single-axis scanning"


def test_waypoint_pose_matches_reference_golden(lrc, golden):
    g = golden("poses.npz")
    for row, ref in zip(g["xyzyaw"], g["pose"]):
        got = lrc.Waypoint(*row).to_pose_matrix()
        assert got.dtype == np.float64 and np.array_equal(got, ref)
    assert lrc.poses_from_waypoints([]).shape == (0, 4, 4)


def test_label_packing_roundtrip(lrc):
    sem = np.array([0, 1, 2, 7, 12, 65535], np.uint16)
    ins = np.array([0, 5, 300, 65535, 1, 2], np.uint16)
    s2, i2 = lrc.unpack_labels(lrc.pack_labels(sem, ins))
    assert np.array_equal(s2, sem) and np.array_equal(i2, ins)


def test_synthetic_meshes_are_deterministic_and_sized(lrc):
    a, b = lrc.synthetic.box_room(), lrc.synthetic.box_room()
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.triangles, b.triangles)
    assert abs(len(a.triangles) - 50_000) <= 500                               # 50k +- 1 %
    assert a.triangles.max() < len(a.vertices) and a.triangles.min() >= 0
    sem, ins = lrc.unpack_labels(a.triangle_labels)
    assert set(np.unique(sem)) == {0, 1, 2, 7, 8, 10}
    wps = lrc.synthetic.office_waypoints(100)
    assert len(wps) == 100 and all(w.z == 1.0 and w.yaw == 0.0 for w in wps)


def test_pinned_result_pool_lends_and_takes_back_buffers(monkeypatch):
    """The per-frame call hands out pooled page-locked buffers AS numpy arrays (core._PinnedPool / _PinnedLease): a buffer
    returns to the pool exactly when the last array built on it is dropped, an exhausted pool says so (the caller then copies
    into ordinary arrays), and arrays stay valid after the pool object itself is gone.  CPU: pinning is patched out."""
    import gc
    import torch
    from lrc_b200 import core
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self, raising=False)
    pool = core._PinnedPool(64, max_items=2)
    a = pool.acquire()
    arr_a = a.array(np.float32, 16)
    arr_a[:] = np.arange(16, dtype=np.float32)
    view = arr_a.reshape(4, 4)[1:]                         # a view keeps the lease alive as well
    del a, arr_a
    b = pool.acquire()
    arr_b = b.array(np.float64, 8)
    arr_b[:] = 7.0
    del b
    assert pool.acquire() is None                          # both buffers are on loan
    assert np.array_equal(view, np.arange(16, dtype=np.float32).reshape(4, 4)[1:]) and np.all(arr_b == 7.0)
    del view
    gc.collect()
    assert len(pool.free) == 1
    c = pool.acquire()                                     # the returned buffer is lent again, no third allocation
    assert c is not None and pool.count == 2 and pool.acquire() is None
    arr_c = c.array(np.float32, 4)
    del c, pool
    gc.collect()
    arr_c[:] = 1.0                                         # still backed by live memory after the pool is gone
    assert arr_c.flags.writeable and arr_c.sum() == 4.0 and np.all(arr_b == 7.0)
