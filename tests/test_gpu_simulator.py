"""GPU: the batched frame loop (lrc_b200.simulator.run_simulation) against the reference's per-waypoint loop
(s3dis_simulator.py:254-294) restated with the oracle engine and the oracle's ScanQuality expressions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class _Lidar:
    def __init__(self, rays, pose, intr):
        self._rays, self.pose, self.intrinsics = rays, pose, intr

    def get_rays(self):
        return self._rays


def test_run_simulation_equals_reference_frame_loop(lrc, orc):
    from oracle import post_oracle as po
    engine = lrc.RaycastEngineGPU(device=0)
    mesh = lrc.synthetic.box_room(target_tris=6000, seed=3)
    wps = [lrc.Waypoint(2.0 + 0.8 * k, 2.5 + 0.3 * k, 1.0, 0.0) for k in range(5)] + [lrc.Waypoint(-50.0, 0.0, 1.0, 0.0)]
    intr = lrc.Indoor8LineLidarIntrinsics(max_range=5.0, horizontal_res=900)
    run = lrc.run_simulation(engine, wps, intr, mesh)
    bounds = lrc.room_bounds_of(mesh)
    vol = lrc.room_volume(bounds)
    assert len(run.frames) == 6 and run.statistics.total_frames == 6
    ref_engine = orc.OracleEngineCPU()
    qs = []
    for i, w in enumerate(wps):                                            # the reference's loop body, :254-288
        pose = w.to_pose_matrix()
        rays = orc.gen_rays_single_axis(pose, intr.vertical_degrees, intr.horizontal_res)
        points, incident = ref_engine.lidar_intersect_mesh(_Lidar(rays, pose, intr), (mesh.vertices, mesh.triangles))
        q = po.scan_quality(points, incident, intr.get_total_points_per_scan(), vol)
        qs.append(q)
        f = run.frames[i]
        assert f.frame_index == i and np.array_equal(f.points, points)     # bit-exact points, same order
        np.testing.assert_allclose(f.incident_angles, incident, rtol=0, atol=1e-9)
        assert f.scan_quality.num_points == q["num_points"] and f.scan_quality.coverage_ratio == q["coverage_ratio"]
        assert f.scan_quality.scan_density == q["scan_density"]
        assert f.scan_quality.incident_angle_mean == pytest.approx(float(q["incident_angle_mean"]), rel=1e-9, abs=1e-9)
        assert f.scan_quality.range_mean == pytest.approx(float(q["range_mean"]), rel=2e-6, abs=1e-12)
        assert np.array_equal(f.labels, mesh.triangle_labels[f.prim_id])
    assert run.frames[5].get_num_points() == 0 and len(run.frames[5].incident_angles) == 0     # sensor far outside: empty frame
    want = po.simulation_stats(qs, run.simulation_time)
    s = run.statistics
    assert s.total_points == want["total_points"] == run.get_total_points()
    assert s.average_coverage == pytest.approx(float(want["average_coverage"]), rel=1e-12)
    assert s.average_incident_angle == pytest.approx(float(want["average_incident_angle"]), rel=1e-9)
    assert s.frames_per_second == pytest.approx(6 / run.simulation_time) and s.frames_per_second > 0      # not the reference's 0 FPS
    assert lrc.run_simulation(engine, [], intr, mesh).statistics.total_frames == 0


def test_example_script_runs(tmp_path):
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "examples", "simulate_room.py"), "--tris", "20000", "--waypoints", "20",
                          "--out", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "frames/s" in res.stdout and os.path.getsize(tmp_path / "combined_pointcloud_with_label.ply") > 1000
