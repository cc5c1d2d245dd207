"""GPU parity tests proper: the CUDA engine, called through the C ABI, against the CPU oracle on the same seeded
inputs and against the committed golden fixtures.  Bar (BASELINE.json north_star): hit masks and triangle ids
bit-exact except grazing-edge ties (<= 1e-5 of rays), |dt| <= 1e-4 m, labels bit-exact for equal ids.  Because the
engine and the oracle share one float32 intersection spec, the tests below demand much more: bit-equal t, ids
and points."""
import numpy as np
import pytest

from conftest import f32_mismatch, ulp_diff

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF
T_TOL = 1e-4          # metres, north_star
TIE_BUDGET = 1e-5     # fraction of rays, north_star


class _Lidar:
    def __init__(self, rays, pose, max_range):
        self._rays, self.pose = rays, pose
        self.intrinsics = type("I", (), {"max_range": max_range})()

    def get_rays(self):
        return self._rays


def _o3d_like(mesh):
    return (mesh.vertices, mesh.triangles)


# ---- ray generation ---------------------------------------------------------------------------------
@pytest.mark.parametrize("preset", ["8line", "32line", "64line", "custom"])
@pytest.mark.parametrize("pose_name", ["identity", "posed"])
def test_get_rays_single_axis_vs_reference_golden(engine, lrc, golden, golden_poses, preset, pose_name):
    g = golden("rays_single_axis.npz")
    intr = lrc.Indoor8LineLidarIntrinsics(vertical_degrees=list(g[f"{preset}/vertical_degrees"]),
                                          horizontal_res=int(g[f"{preset}/W"]), vertical_res=len(g[f"{preset}/vertical_degrees"]))
    rays = lrc.IndoorLidar(intr, golden_poses[pose_name]).get_rays()
    assert rays.dtype == np.float32 and rays.shape == (int(g[f"{preset}/{pose_name}/n"]), 6)
    sample = g[f"{preset}/{pose_name}/sample"]
    assert not f32_mismatch(rays[::37], sample).any()   # device libm vs numpy: at most the last float32 bit ...
    assert (ulp_diff(rays[::37], sample) > 0).mean() <= 1e-3   # ... and almost always not even that
    np.testing.assert_allclose(rays.astype(np.float64).sum(0), g[f"{preset}/{pose_name}/sum"], atol=1e-3)


def test_get_rays_uniform_fov_vs_reference_golden(engine, lrc, golden, golden_poses):
    g = golden("rays_uniform.npz")
    for key in g.files:
        hw, pose_name, _ = key.split("/")
        H, W = map(int, hw.split("x"))
        intr = lrc.Indoor8LineLidarIntrinsics(vertical_res=H, horizontal_res=W, vertical_degrees=None)
        rays = lrc.IndoorLidar(intr, golden_poses[pose_name]).get_rays()
        assert not f32_mismatch(rays, g[key]).any(), key


def test_get_rays_dual_axis_vs_reference_golden(engine, lrc, golden, golden_poses):
    g = golden("rays_dual_axis.npz")
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    intr.angle_noise_std = 0.0
    intr.dropout_probability = 0.0
    for pose_name, pose in golden_poses.items():
        rays = lrc.DualAxisLidar(intr, pose, seed=1).get_rays()
        assert rays.shape == (64000, 6)
        sample = g[f"blk2go/{pose_name}/sample"]
        assert not f32_mismatch(rays[::37], sample).any()
        assert (ulp_diff(rays[::37], sample) > 0).mean() <= 1e-3
    small = lrc.DualAxisLidarIntrinsics(point_rate=1000, scan_duration=0.5, num_vertical_lines=7, swing_frequency=3.0,
                                        swing_amplitude=0.3, angle_noise_std=0.0, dropout_probability=0.0)
    rays = lrc.DualAxisLidar(small, golden_poses["posed"], seed=1).get_rays()
    assert not f32_mismatch(rays, g["small/rays"]).any()


def test_dual_axis_noise_matches_oracle_philox(engine, lrc, orc, golden_poses):
    """Same counter-based stream on both sides: kept-ray masks are identical, directions agree to float32 rounding."""
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=0xC0FFEE, pose_index_base=41)
    rays, keep = engine.ctx.gen_rays(golden_poses["posed"][None], intr, noise)
    rays, keep = rays.cpu().numpy(), keep.cpu().numpy().astype(bool)
    ref, ref_keep = orc.gen_rays_dual_axis(golden_poses["posed"], orc.dual_params(intr), seed=0xC0FFEE, pose_idx=41, compact=False)
    assert np.array_equal(keep, ref_keep)
    assert abs(keep.mean() - 0.98) < 0.003
    assert not f32_mismatch(rays, ref).any()                      # Box-Muller through two different libms
    assert (ulp_diff(rays, ref) > 0).mean() <= 1e-2
    lid = lrc.DualAxisLidar(intr, golden_poses["posed"], seed=0xC0FFEE, frame_index=41)
    assert np.array_equal(lid.get_rays(), rays[keep])


# ---- intersection -------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c1(lrc, orc):
    mesh = lrc.synthetic.box_room()
    pose = lrc.synthetic.box_room_pose()
    intr = lrc.Indoor8LineLidarIntrinsics.create_standard_8line()
    rays = orc.gen_rays_single_axis(pose, intr.vertical_degrees, intr.horizontal_res)
    scene = orc.OracleScene(_o3d_like(mesh))
    t, pid = scene.cast_rays(rays)
    return dict(mesh=mesh, pose=pose, intr=intr, rays=rays, t=t, pid=pid)


def test_c1_cast_rays_bit_exact(engine, c1):
    """BASELINE config 1: one 8-line frame on the ~50k-triangle box room."""
    t, pid = engine.cast_rays(c1["rays"], c1["mesh"])
    mism = (pid != c1["pid"])
    assert mism.mean() <= TIE_BUDGET
    assert mism.sum() == 0, "expected bit-exact ids on this scene"
    assert np.array_equal(t, c1["t"])
    both = (pid != MISS) & (c1["pid"] != MISS)
    assert np.abs(t[both] - c1["t"][both]).max() <= T_TOL
    info = engine.ctx.bvh_info()
    assert info["num_tris"] == len(c1["mesh"].triangles) and info["max_depth"] < 64


def test_c1_gpu_bruteforce_equals_bvh(engine, c1):
    engine.set_mesh(c1["mesh"])
    t_bvh, p_bvh = engine.ctx.cast_rays(c1["rays"])
    t_bf, p_bf = engine.ctx.cast_rays(c1["rays"], bruteforce=True)
    assert bool((t_bvh == t_bf).all()) and bool((p_bvh == p_bf).all())


def test_c1_lidar_intersect_mesh_equals_oracle_engine(engine, lrc, orc, c1):
    pts, inc = engine.lidar_intersect_mesh(lrc.create_lidar(c1["intr"], c1["pose"]), c1["mesh"])
    ref = orc.OracleEngineCPU()
    rpts, rinc = ref.lidar_intersect_mesh(_Lidar(c1["rays"], c1["pose"], c1["intr"].max_range), _o3d_like(c1["mesh"]))
    assert pts.dtype == np.float32 and inc.dtype == np.float64
    assert np.array_equal(pts, rpts)
    np.testing.assert_allclose(inc, rinc, rtol=0, atol=1e-9)
    scan = engine.last_scan.numpy()
    assert np.array_equal(scan["prim_id"], ref.last_prim_id)
    assert np.array_equal(scan["label"], c1["mesh"].triangle_labels[scan["prim_id"]])
    assert np.all(np.diff(scan["ray_idx"].astype(np.int64)) > 0)


def test_counters_report_work(engine, c1):
    engine.set_mesh(c1["mesh"])
    engine.ctx.set_counting(True)
    try:
        engine.ctx.counters(reset=True)
        engine.ctx.cast_rays(c1["rays"])
        c = engine.ctx.counters(reset=True)
    finally:
        engine.ctx.set_counting(False)
    assert c["rays"] == len(c1["rays"]) and c["hits"] == int((c1["pid"] != MISS).sum())
    assert 10 < c["nodes_visited"] / c["rays"] < 400 and 1 <= c["tris_tested"] / c["rays"] < 100


# ---- golden frames produced by the reference's own engine code -----------------------------------------
@pytest.fixture(scope="module")
def frame(golden):
    return golden("frame_box_room.npz")


def test_golden_frames_through_reference_shaped_api(engine, lrc, frame, golden_poses):
    mesh = lrc.TriangleMesh(frame["verts"].astype(np.float64), frame["tris"], frame["labels"])
    pose = golden_poses["posed"]
    pts, inc = engine.lidar_intersect_mesh(lrc.create_lidar(lrc.Indoor8LineLidarIntrinsics.create_standard_8line(), pose), mesh)
    assert np.array_equal(pts, frame["8line/points"])
    np.testing.assert_allclose(inc, frame["8line/incident"], rtol=0, atol=1e-9)
    short = lrc.Indoor8LineLidarIntrinsics(max_range=3.0, horizontal_res=500)
    pts, inc = engine.lidar_intersect_mesh(lrc.create_lidar(short, pose), mesh)
    assert np.array_equal(pts, frame["short/points"])
    np.testing.assert_allclose(inc, frame["short/incident"], rtol=0, atol=1e-9)
    # explicit rays with misses: ordered compaction
    pts = engine.rays_intersect_mesh(frame["outside/rays"], mesh)
    assert np.array_equal(pts, frame["outside/points"])
    out_pose = lrc.Waypoint(-3.0, 4.0, 1.0, 0.0).to_pose_matrix()
    pts, inc = engine.lidar_intersect_mesh(_Lidar(frame["outside/rays"], out_pose, 20.0), mesh)    # duck-typed sensor
    assert np.array_equal(pts, frame["outside/lidar_points"])
    np.testing.assert_allclose(inc, frame["outside/lidar_incident"], rtol=0, atol=1e-9)
    pts2, inc2 = engine.lidar_intersect_mesh(lrc.create_lidar(lrc.Indoor8LineLidarIntrinsics(horizontal_res=400), out_pose), mesh)
    assert np.array_equal(pts2, pts) and np.array_equal(inc2, inc)                                  # fused path == explicit path


# ---- trajectories ----------------------------------------------------------------------------------
def test_simulate_equals_per_frame_calls_and_oracle(engine, lrc, orc, c1):
    poses = lrc.poses_from_waypoints([lrc.Waypoint(2.0 + 0.7 * k, 3.0 + 0.2 * k, 1.0, 0.1 * k) for k in range(5)])
    intr = lrc.Indoor8LineLidarIntrinsics(horizontal_res=720, max_range=6.0)
    res = engine.simulate(poses, intr, c1["mesh"])
    assert res.num_frames == 5
    scene = orc.OracleScene(_o3d_like(c1["mesh"]))
    for p in range(5):
        pts, inc = res.frame(p)
        one_pts, one_inc = engine.lidar_intersect_mesh(lrc.create_lidar(intr, poses[p]), c1["mesh"])
        assert np.array_equal(pts, one_pts) and np.array_equal(inc, one_inc)
        rays = orc.gen_rays_single_axis(poses[p], intr.vertical_degrees, intr.horizontal_res)
        fr = orc.epilogue_c(rays, *scene.cast_rays(rays), center=poses[p][:3, 3], max_range=intr.max_range)
        assert np.array_equal(pts, fr.points)
        np.testing.assert_allclose(inc, fr.incident, rtol=0, atol=1e-9)
    # chunked execution (bounded scratch) must not change a bit
    full = res.numpy()
    engine.ctx.set_option("chunk_rays", 720 * 8)        # one frame per chunk
    try:
        chunked = engine.simulate(poses, intr).numpy()
    finally:
        engine.ctx.set_option("chunk_rays", 1 << 26)
    for k in full:
        assert np.array_equal(full[k], chunked[k]), k


def test_dual_axis_trajectory_with_noise_and_labels(engine, lrc, orc, c1):
    """BASELINE config 3 in miniature: BLK2GO pattern, angle noise + dropout through Philox, labels gathered."""
    poses = lrc.poses_from_waypoints([lrc.Waypoint(3.0 + k, 3.5, 1.0, 0.0) for k in range(3)])
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2, pose_index_base=10)
    res = engine.simulate(poses, intr, c1["mesh"], noise=noise)
    out = res.numpy()
    scene = orc.OracleScene(_o3d_like(c1["mesh"]))
    labels = c1["mesh"].triangle_labels
    total, bad = 0, 0
    for p in range(3):
        rays, keep = orc.gen_rays_dual_axis(poses[p], orc.dual_params(intr), seed=2, pose_idx=10 + p, compact=False)
        t, pid = scene.cast_rays(rays)
        fr = orc.epilogue_c(rays, t, pid, center=poses[p][:3, 3], max_range=intr.max_range, tri_label=labels, keep=keep.astype(np.uint8))
        a, b = out["frame_offset"][p], out["frame_offset"][p + 1]
        # compare ray by ray via ray_idx; the id budget is the north-star's grazing-edge tie rate (1e-5 of rays)
        assert np.array_equal(out["ray_idx"][a:b], fr.ray_idx)          # same rays survive dropout / hit / range
        same = out["prim_id"][a:b] == fr.prim_id
        total += len(same)
        bad += int((~same).sum())
        assert np.abs(out["points"][a:b][same] - fr.points[same]).max() <= T_TOL
        assert np.array_equal(out["label"][a:b], labels[out["prim_id"][a:b]])
        assert abs(keep.mean() - 0.98) < 0.004
    assert bad / total <= 1e-5, (bad, total)


# ---- edge cases -----------------------------------------------------------------------------------
def test_tiny_and_degenerate_meshes(engine, lrc, orc):
    rays = np.array([[0.25, 0.25, 0, 0, 0, 1], [0.25, 0.25, 0, 0, 0, -1], [5, 5, 0, 0, 0, 1]], np.float32)
    one = (np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], float), np.array([[0, 1, 2]], np.int32))
    t, pid = engine.cast_rays(rays, one)
    assert list(pid) == [0, MISS, MISS] and t[0] == 1.0 and np.isinf(t[1])
    two = (np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1], [0, 0, 2], [1, 0, 2], [0, 1, 2]], float), np.array([[3, 4, 5], [0, 1, 2]], np.int32))
    t, pid = engine.cast_rays(rays, two)
    assert list(pid) == [1, MISS, MISS] and t[0] == 1.0
    # coincident triangles: the tie goes to the smallest id
    tri = [[0, 0, 1], [1, 0, 1], [0, 1, 1]]
    dup = (np.array(tri * 3, float), np.arange(9, dtype=np.int32).reshape(3, 3)[::-1].copy())
    t, pid = engine.cast_rays(rays, dup)
    assert pid[0] == 0 and t[0] == 1.0
    # degenerate triangles never hit, and do not break the build
    deg = (np.array([[0, 0, 1], [1, 1, 1], [2, 2, 1], [0, 0, 2], [0, 0, 2], [0, 0, 2]], float), np.array([[0, 1, 2], [3, 4, 5]], np.int32))
    t, pid = engine.cast_rays(rays, deg)
    assert np.all(pid == MISS)
    # empty mesh and empty ray set
    empty = (np.zeros((0, 3)), np.zeros((0, 3), np.int32))
    t, pid = engine.cast_rays(rays, empty)
    assert np.all(pid == MISS) and np.all(np.isinf(t))
    assert engine.rays_intersect_mesh(rays, empty).shape == (0, 3)
    assert engine.rays_intersect_mesh(np.zeros((0, 6), np.float32), one).shape == (0, 3)
    pts, inc = engine.lidar_intersect_mesh(_Lidar(np.zeros((0, 6), np.float32), np.eye(4), 10.0), one)
    assert pts.shape == (0, 3) and inc.shape == (0,)


def test_input_checks_match_reference(engine, lrc):
    one = (np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], float), np.array([[0, 1, 2]], np.int32))
    with pytest.raises(TypeError):
        engine.rays_intersect_mesh([[0, 0, 0, 0, 0, 1]], one)             # reference raycast_engine_cpu.py:40-41
    with pytest.raises(ValueError):
        engine.rays_intersect_mesh(np.zeros((3, 5), np.float32), one)     # reference :42-43
    bad = (np.zeros((3, 3)), np.array([[0, 1, 7]], np.int32))
    with pytest.raises(RuntimeError, match="outside"):
        engine.set_mesh(bad)                                              # fails loudly, unlike s3dis_simulator.py:271-273


def test_documented_edge_leak_is_reproduced_bit_for_bit(engine, lrc, orc):
    mesh = lrc.synthetic.empty_box((0, 0, 0), (4, 3, 2.5), pitch=0.5)
    ray = np.array([[1.3, 0.9, 1.1, 4.3297803e-17, 0.70710677, -0.70710677]], np.float32)
    t, pid = engine.cast_rays(ray, mesh)
    tb, pb = orc.cast_rays_brute(_o3d_like(mesh), ray)
    assert pid[0] == pb[0] and t[0] == tb[0]


def test_unjittered_box_edges_and_corners(engine, lrc, orc):
    """Axis-aligned, un-jittered geometry with rays through vertices and along faces: every outcome (hit, tie,
    leak) must equal the oracle's."""
    mesh = lrc.synthetic.empty_box((0, 0, 0), (4, 3, 2.5), pitch=0.5)
    pose = np.eye(4)
    pose[:3, 3] = (2.0, 1.5, 1.0)                        # on grid planes: many rays hit edges/vertices exactly
    rays = orc.gen_rays_single_axis(pose, [45.0, 0.0, -45.0, -90.0, 90.0], 720)
    t, pid = engine.cast_rays(rays, mesh)
    rt, rpid = orc.OracleScene(_o3d_like(mesh)).cast_rays(rays)
    bt, bpid = orc.cast_rays_brute(_o3d_like(mesh), rays)
    assert np.array_equal(rpid, bpid) and np.array_equal(rt, bt)
    assert np.array_equal(pid, rpid) and np.array_equal(t, rt)


def test_simulate_to_host_pipelined_equals_simulate(engine, lrc, c1):
    """The host-output path (pose chunks, D2H overlapped with later chunks' kernels) returns the same bits."""
    poses = lrc.poses_from_waypoints([lrc.Waypoint(2.0 + 0.5 * k, 3.0 + 0.2 * k, 1.0, 0.1 * k) for k in range(7)])
    for intr, noise in ((lrc.Indoor8LineLidarIntrinsics(horizontal_res=720, max_range=6.0), None),
                        (lrc.DualAxisLidarIntrinsics(point_rate=64000, scan_duration=0.1, max_range=5.0),
                         lrc.NoiseConfig(0.001, 0.02, 0.0, seed=5, pose_index_base=3))):
        ref = engine.simulate(poses, intr, c1["mesh"], noise=noise).numpy()
        for chunk in (1, 2, 7, None):
            got = engine.simulate_to_host(poses, intr, noise=noise, chunk_poses=chunk)
            assert got["num_points"] == len(ref["points"]) > 1000
            assert np.array_equal(got["frame_offset"], ref["frame_offset"])
            assert np.array_equal(got["points"], ref["points"]) and np.array_equal(got["incident"], ref["incident"])
            assert np.array_equal(got["label"], ref["label"])


def test_gather_targets_and_pipelined_compaction_single_gpu(engine, lrc, c1):
    """lrc_set_gather with this GPU's own buffer as the only target: the fused store path and the two-stream
    chunk pipeline (compaction of chunk c behind the traversal of chunk c+1) must not change a bit."""
    from lrc_b200.distributed import PeerGather
    poses = lrc.poses_from_waypoints([lrc.Waypoint(2.0 + 0.5 * k, 3.0 + 0.2 * k, 1.0, 0.1 * k) for k in range(7)])
    intr = lrc.Indoor8LineLidarIntrinsics(horizontal_res=720, max_range=6.0)
    ref = engine.simulate(poses, intr, c1["mesh"]).numpy()
    pg = PeerGather(engine.ctx, cap_per_rank=7 * 8 * 720, frames_per_rank=7)
    try:
        for chunks in (1, 2, 3, 7):
            engine.ctx.set_option("gather_chunks", chunks)
            pg.enable()
            got_local = engine.simulate(poses, intr).numpy()
            pg.synchronize()
            got = pg.assemble_numpy()
            pg.disable()
            for k in ref:
                assert np.array_equal(got_local[k], ref[k]), (chunks, k)
            assert np.array_equal(got["frame_offset"], ref["frame_offset"])
            assert np.array_equal(got["points"], ref["points"]) and np.array_equal(got["label"], ref["label"])
    finally:
        engine.ctx.set_option("gather_chunks", 4)
        pg.close()


def test_all_traversal_variants_and_block_sizes_give_identical_bits(engine, lrc, orc, c1):
    """Loop shape (if-if / while-while), 256-bit node loads, the 32-register build, the shared-memory top of the tree
    (1..8 levels, also on trees smaller than the table) and the block size are tuning knobs: none may change a bit."""
    ctx = engine.ctx
    poses = lrc.poses_from_waypoints([lrc.Waypoint(3.1, 2.7, 1.0, 0.3), lrc.Waypoint(7.0, 5.0, 1.2, 2.0)])
    intr = lrc.Indoor8LineLidarIntrinsics(max_range=6.0, horizontal_res=1000)
    tiny = lrc.TriangleMesh(np.array([[0, -5, -5], [0, 5, -5], [0, 0, 5.0], [9, -5, -5], [9, 5, -5], [9, 0, 5.0]]) + [4.0, 3, 1], [[0, 1, 2], [3, 4, 5]])
    try:
        for mesh in (c1["mesh"], tiny, lrc.TriangleMesh(tiny.vertices[:3], [[0, 1, 2]])):
            ctx.set_option("variant", 1); ctx.set_option("block", 128)
            ref = engine.simulate(poses, intr, mesh).numpy()
            for var, blk, top in ((0, 128, 0), (5, 64, 0), (2, 32, 0), (3, 128, 0), (13, 128, 1), (13, 64, 4), (13, 128, 8),
                                  (21, 128, -1), (21, 64, -3), (21, 128, -40)):
                ctx.set_option("variant", var); ctx.set_option("block", blk)
                if top > 0:
                    ctx.set_option("top_levels", top)
                if top < 0:
                    ctx.set_option("stack_levels", -top)
                got = engine.simulate(poses, intr, mesh).numpy()
                for k in ref:
                    assert np.array_equal(got[k], ref[k]), (var, blk, top, k)
            # 32-byte quantised node records: a different tree encoding, the same bits out
            ctx.set_option("node_format", 1); ctx.invalidate_mesh()
            try:
                got = engine.simulate(poses, intr, mesh).numpy()
                assert engine.ctx.bvh_info()["bytes_nodes"] == 32 * engine.ctx.bvh_info()["num_nodes"]
            finally:
                ctx.set_option("node_format", 0); ctx.invalidate_mesh()
            for k in ref:
                assert np.array_equal(got[k], ref[k]), ("node_format 1", k)
            # multi-triangle leaves (2, 4, 8 Morton-consecutive triangles per leaf), under both record formats
            for leaf, fmt in ((1, 0), (4, 0), (8, 0), (1, 1), (4, 1)):
                ctx.set_option("leaf_size", leaf); ctx.set_option("node_format", fmt); ctx.invalidate_mesh()
                try:
                    got = engine.simulate(poses, intr, mesh).numpy()
                finally:
                    ctx.set_option("leaf_size", 2); ctx.set_option("node_format", 0); ctx.invalidate_mesh()
                for k in ref:
                    assert np.array_equal(got[k], ref[k]), ("leaf_size", leaf, fmt, k)
    finally:      # back to the library's defaults (paired records, pop culling, block size by call)
        ctx.set_option("variant", 65); ctx.set_option("block", 0); ctx.set_option("top_levels", 6); ctx.set_option("stack_levels", 12)
        ctx.set_option("node_format", 2); ctx.set_option("leaf_size", 2); ctx.invalidate_mesh()


def test_round2_kernel_options_give_identical_bits(engine, lrc, c1):
    """Everything round 2 added on top of the paired-record kernel is a tuning knob: prefetch / barrier-free counts / streaming
    stores / texture-pipe loads (``tune``), warp packets, K rays per thread, persistent warps, pose chunks on one GPU, block
    size by call, the PLOC builder and node compaction.  None may change a bit -- single-axis and noisy dual-axis scans."""
    ctx = engine.ctx
    poses = lrc.poses_from_waypoints([lrc.Waypoint(3.1 + 0.4 * k, 2.7 + 0.1 * k, 1.0, 0.3 * k) for k in range(5)])
    dual = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    cases = ((lrc.Indoor8LineLidarIntrinsics(max_range=6.0, horizontal_res=1003), None),
             (dual, lrc.NoiseConfig.from_intrinsics(dual, seed=3, pose_index_base=7)))
    defaults = {"tune": 2, "warp_packet": 0, "rays_per_thread": 1, "persistent": 0, "scan_chunks": 1, "scan_taper": 1, "block": 0,
                "build_quality": 0, "compact_nodes": 0}
    variants = [{"tune": t} for t in (0, 1, 4, 6, 7, 10, 18, 26)] + [{"warp_packet": 1}, {"rays_per_thread": 2}, {"rays_per_thread": 4},
                {"persistent": 1}, {"persistent": 2}, {"scan_chunks": 3}, {"scan_chunks": 2, "scan_taper": 3}, {"block": 32}, {"block": 64},
                {"block": 128}, {"build_quality": 1}, {"compact_nodes": 1}, {"build_quality": 1, "compact_nodes": 1, "warp_packet": 1}]
    try:
        ctx.set_option("node_format", 2); ctx.set_option("variant", 65)
        for k, v in defaults.items():
            ctx.set_option(k, v)
        ctx.invalidate_mesh()
        refs = [engine.simulate(poses, intr, c1["mesh"], noise=nz).numpy() for intr, nz in cases]
        assert all(r["frame_offset"][-1] > 10000 for r in refs)
        for var in variants:
            for k, v in var.items():
                ctx.set_option(k, v)
            if "build_quality" in var or "compact_nodes" in var:
                ctx.invalidate_mesh()
            try:
                # scan_chunks only cuts calls of >= 2^20 rays per chunk: make the threshold reachable with 5 frames of 64000 rays? it is
                # not -- the option must then simply have no effect, which is what this asserts for small calls
                for (intr, nz), ref in zip(cases, refs):
                    got = engine.simulate(poses, intr, c1["mesh"], noise=nz).numpy()
                    for key in ref:
                        assert np.array_equal(got[key], ref[key]), (var, type(intr).__name__, key)
            finally:
                for k in var:
                    ctx.set_option(k, defaults[k])
                if "build_quality" in var or "compact_nodes" in var:
                    ctx.invalidate_mesh()
    finally:
        for k, v in defaults.items():
            ctx.set_option(k, v)
        ctx.invalidate_mesh()
