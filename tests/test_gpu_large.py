"""Parity at BASELINE.json's full sizes (1M-triangle office): oracle comparison on a pose subset, plus
size-independent properties over the whole trajectory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF


@pytest.fixture(scope="module")
def office(lrc):
    return lrc.synthetic.office()


@pytest.fixture(scope="module")
def office_poses(lrc):
    return lrc.poses_from_waypoints(lrc.synthetic.office_waypoints(100))


def test_c2_subset_bit_exact_vs_oracle(engine, lrc, orc, office, office_poses):
    """BASELINE config 2 (32-line, 1M triangles): ALL 100 poses (12.8 M rays) against the CPU oracle's SAH BVH."""
    intr = lrc.Indoor8LineLidarIntrinsics.create_dense_32line()
    sel = office_poses
    res = engine.simulate(sel, intr, office).numpy()
    info = engine.ctx.bvh_info()
    assert 990_000 <= info["num_tris"] <= 1_010_000 and info["max_depth"] < 64
    scene = orc.OracleScene((office.vertices, office.triangles))
    n_bad, n_all = 0, 0
    for k in range(len(sel)):
        rays = orc.gen_rays_single_axis(sel[k], intr.vertical_degrees, intr.horizontal_res)
        t, pid = scene.cast_rays(rays)
        fr = orc.epilogue_c(rays, t, pid, center=sel[k][:3, 3], max_range=intr.max_range, tri_label=office.triangle_labels)
        a, b = res["frame_offset"][k], res["frame_offset"][k + 1]
        assert b - a == len(fr.points)
        assert np.array_equal(res["ray_idx"][a:b], fr.ray_idx)
        same = res["prim_id"][a:b] == fr.prim_id
        n_bad += int((~same).sum())
        n_all += len(same)
        assert np.array_equal(res["points"][a:b][same], fr.points[same])
        assert np.abs(res["points"][a:b] - fr.points).max() <= 1e-4
        np.testing.assert_allclose(res["incident"][a:b][same], fr.incident[same], rtol=0, atol=1e-9)
        assert np.array_equal(res["label"][a:b], office.triangle_labels[res["prim_id"][a:b]])
    assert n_bad / n_all <= 1e-5


def test_c2_full_trajectory_properties(engine, lrc, office, office_poses):
    """All 100 poses x 128000 rays: closed room => every ray hits; hits lie inside the room; frames are
    ray-ordered; running the trajectory as two halves (the multi-GPU sharding) concatenates to the same bits."""
    intr = lrc.Indoor8LineLidarIntrinsics.create_dense_32line()
    res = engine.simulate(office_poses, intr, office)
    out = res.numpy()
    off = out["frame_offset"]
    assert res.num_frames == 100 and off[0] == 0 and np.all(np.diff(off) > 0)
    assert off[-1] >= 0.999 * 100 * 128000                      # closed shell, 25 m range: (almost) every ray returns
    p = out["points"]
    assert p[:, 0].min() > -1e-2 and p[:, 0].max() < 20.01 and p[:, 1].min() > -1e-2 and p[:, 1].max() < 15.01
    assert p[:, 2].min() > -1e-2 and p[:, 2].max() < 3.01
    for f in (0, 50, 99):
        r = out["ray_idx"][off[f]:off[f + 1]].astype(np.int64)
        assert np.all(np.diff(r) > 0) and r.max() < 128000
    assert np.array_equal(out["label"], office.triangle_labels[out["prim_id"]])
    # pose chunks on one GPU (compaction of chunk c behind the traversal of c+1; option, default off) and the warp-packet kernel
    for opts in ({"scan_chunks": 3, "scan_taper": 3}, {"warp_packet": 1}):
        for k, v in opts.items():
            engine.ctx.set_option(k, v)
        try:
            alt = engine.simulate(office_poses, intr).numpy()
        finally:
            for k in opts:
                engine.ctx.set_option(k, 1 if k.startswith("scan") else 0)
        for k in ("points", "incident", "prim_id", "label", "ray_idx", "frame_offset"):
            assert np.array_equal(alt[k], out[k]), (opts, k)
    a = engine.simulate(office_poses[:50], intr).numpy()
    b = engine.simulate(office_poses[50:], intr).numpy()
    for k in ("points", "incident", "prim_id", "label", "ray_idx"):
        assert np.array_equal(np.concatenate([a[k], b[k]]), out[k]), k
    assert np.array_equal(np.concatenate([a["frame_offset"], a["frame_offset"][-1] + b["frame_offset"][1:]]), off)


def test_c3_blk2go_noise_labels_subset_vs_oracle(engine, lrc, orc, office, office_poses):
    """BASELINE config 3 (dual-axis BLK2GO, angle noise + dropout + labels, 1M triangles): 64 consecutive poses (4 M
    rays) against the oracle's identical Philox stream; dropped rays never reach the output, labels follow triangle ids."""
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2, pose_index_base=30)
    sel = office_poses[30:94]
    res = engine.simulate(sel, intr, office, noise=noise).numpy()
    scene = orc.OracleScene((office.vertices, office.triangles))
    n_bad = n_all = 0
    for k in range(len(sel)):
        rays, keep = orc.gen_rays_dual_axis(sel[k], orc.dual_params(intr), seed=2, pose_idx=30 + k, compact=False)
        kept = np.nonzero(keep)[0]
        assert 0.97 * 64000 < len(kept) < 0.99 * 64000                    # ~2 % dropout
        t, pid = scene.cast_rays(rays[kept])
        fr = orc.epilogue_c(rays[kept], t, pid, center=sel[k][:3, 3], max_range=intr.max_range, tri_label=office.triangle_labels)
        a, b = res["frame_offset"][k], res["frame_offset"][k + 1]
        assert b - a == len(fr.points)
        assert np.array_equal(res["ray_idx"][a:b], kept[fr.ray_idx])       # indices are positions in the DENSE table
        same = res["prim_id"][a:b] == fr.prim_id
        n_bad += int((~same).sum())
        n_all += len(same)
        assert np.array_equal(res["points"][a:b][same], fr.points[same])
        assert np.array_equal(res["label"][a:b], office.triangle_labels[res["prim_id"][a:b]])
    assert n_bad / n_all <= 1e-5


def test_c4_floor_chunking_and_host_path_properties(engine, lrc):
    """BASELINE config 4's mesh (multi-room floor, 5M triangles): the BVH does not fit L2.  Size-independent
    properties: tiny scratch chunks (forces the pipelined multi-chunk path), the host-buffer path and the one-launch
    path produce identical bits; re-running is idempotent; frames are ray-ordered and labelled by triangle."""
    mesh = lrc.synthetic.floor_plan(target_tris=5_000_000, seed=0)
    poses = lrc.poses_from_waypoints(lrc.synthetic.floor_plan_waypoints(24))
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=9)
    ref = engine.simulate(poses, intr, mesh, noise=noise).numpy()
    info = engine.ctx.bvh_info()
    assert 4_900_000 <= info["num_tris"] <= 5_100_000 and info["max_depth"] < 64
    assert info["bytes_nodes"] + info["bytes_tris"] > 4 * 126e6
    again = engine.simulate(poses, intr, noise=noise).numpy()
    engine.ctx.set_option("chunk_rays", 5 * 64000)                          # 5 poses per chunk -> 5 chunks, double-buffered
    try:
        chunked = engine.simulate(poses, intr, noise=noise).numpy()
    finally:
        engine.ctx.set_option("chunk_rays", 1 << 26)
    host = engine.simulate_to_host(poses, intr, noise=noise, chunk_poses=7)
    for k in ("points", "incident", "prim_id", "label", "ray_idx", "frame_offset"):
        assert np.array_equal(again[k], ref[k]), k
        assert np.array_equal(chunked[k], ref[k]), k
    assert np.array_equal(host["points"], ref["points"]) and np.array_equal(host["incident"], ref["incident"])
    assert np.array_equal(host["label"], ref["label"]) and np.array_equal(host["frame_offset"], ref["frame_offset"])
    off = ref["frame_offset"]
    assert off[-1] > 0.9 * 24 * 64000
    for f in (0, 11, 23):
        r = ref["ray_idx"][off[f]:off[f + 1]].astype(np.int64)
        assert np.all(np.diff(r) > 0)
    assert np.array_equal(ref["label"], mesh.triangle_labels[ref["prim_id"]])
    # a different seed changes the cloud, the same seed with a shifted pose base reproduces the tail of the trajectory
    tail = engine.simulate(poses[12:], intr, noise=lrc.NoiseConfig.from_intrinsics(intr, seed=9, pose_index_base=12)).numpy()
    assert np.array_equal(tail["points"], ref["points"][off[12]:])
    other = engine.simulate(poses[:2], intr, noise=lrc.NoiseConfig.from_intrinsics(intr, seed=10)).numpy()
    assert not np.array_equal(other["frame_offset"], ref["frame_offset"][:3]) or not np.array_equal(other["points"], ref["points"][:off[2]])


def _compare_single_axis(orc, scene, mesh, res, poses, intr):
    """Frames of ``res`` against the oracle, frame by frame; returns (mismatching ids, rays compared)."""
    n_bad = n_all = 0
    for k in range(len(poses)):
        rays = orc.gen_rays_single_axis(poses[k], intr.vertical_degrees, intr.horizontal_res)
        t, pid = scene.cast_rays(rays)
        fr = orc.epilogue_c(rays, t, pid, center=poses[k][:3, 3], max_range=intr.max_range, tri_label=mesh.triangle_labels)
        a, b = res["frame_offset"][k], res["frame_offset"][k + 1]
        assert b - a == len(fr.points), k
        assert np.array_equal(res["ray_idx"][a:b], fr.ray_idx), k
        same = res["prim_id"][a:b] == fr.prim_id
        n_bad += int((~same).sum())
        n_all += len(same)
        assert np.array_equal(res["points"][a:b][same], fr.points[same]), k          # bit-exact where ids match
        assert np.abs(res["points"][a:b] - fr.points).max() <= 1e-4                     # north-star: |dt| <= 1e-4 m
        np.testing.assert_allclose(res["incident"][a:b][same], fr.incident[same], rtol=0, atol=1e-9)
        assert np.array_equal(res["label"][a:b], mesh.triangle_labels[res["prim_id"][a:b]]), k
    return n_bad, n_all


def test_c4_floor_subset_bit_exact_vs_oracle(engine, lrc, orc):
    """BASELINE config 4 (BLK2GO + noise + labels on the 5M-triangle multi-room floor, BVH 560 MB > L2): 16 poses spread
    over the 500-pose coverage trajectory against the oracle's SAH BVH and its identical Philox stream."""
    mesh = lrc.synthetic.floor_plan(target_tris=5_000_000, seed=0)
    poses_all = lrc.poses_from_waypoints(lrc.synthetic.floor_plan_waypoints(500))
    idx = list(range(0, 500, 34)) + [499]
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    scene = orc.OracleScene((mesh.vertices, mesh.triangles))
    n_bad = n_all = 0
    first = True
    for gi in idx:
        noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2, pose_index_base=gi)
        res = engine.simulate(poses_all[gi:gi + 1], intr, mesh if first else None, noise=noise).numpy()
        first = False
        rays, keep = orc.gen_rays_dual_axis(poses_all[gi], orc.dual_params(intr), seed=2, pose_idx=gi, compact=False)
        kept = np.nonzero(keep)[0]
        t, pid = scene.cast_rays(rays[kept])
        fr = orc.epilogue_c(rays[kept], t, pid, center=poses_all[gi][:3, 3], max_range=intr.max_range, tri_label=mesh.triangle_labels)
        assert res["frame_offset"][1] == len(fr.points) and len(fr.points) > 30000
        assert np.array_equal(res["ray_idx"], kept[fr.ray_idx])
        same = res["prim_id"] == fr.prim_id
        n_bad += int((~same).sum())
        n_all += len(same)
        assert np.array_equal(res["points"][same], fr.points[same])
        assert np.abs(res["points"] - fr.points).max() <= 1e-4
        np.testing.assert_allclose(res["incident"][same], fr.incident[same], rtol=0, atol=1e-9)
        assert np.array_equal(res["label"], mesh.triangle_labels[res["prim_id"]])
    info = engine.ctx.bvh_info()
    assert 4_900_000 <= info["num_tris"] <= 5_100_000
    assert n_bad / n_all <= 1e-5, (n_bad, n_all)


@pytest.mark.parametrize("target_tris", [10_000, 100_000, 10_000_000])
def test_c5_sweep_points_bit_exact_vs_oracle(engine, lrc, orc, target_tris):
    """BASELINE config 5's mesh-size axis (1e4 ... 1e7 triangles; 1e6 is the C2 test above): 32-line frames on the
    synthetic office against the oracle.  At 1e7 triangles (BVH 1.1 GB) the trajectory is also run through the
    multi-chunk path the 1e9-ray sweep uses (``chunk_rays`` forced to two frames) and must give the single-chunk bits."""
    mesh = lrc.synthetic.office(target_tris=target_tris, seed=0)
    intr = lrc.Indoor8LineLidarIntrinsics.create_dense_32line()
    poses = lrc.poses_from_waypoints(lrc.synthetic.office_waypoints(100))[[3, 48, 77, 91, 12]]
    n_pose = 5 if target_tris >= 10_000_000 else 2
    poses = poses[:n_pose]
    res = engine.simulate(poses, intr, mesh).numpy()
    info = engine.ctx.bvh_info()
    assert 0.9 * target_tris <= info["num_tris"] <= 1.1 * target_tris and info["max_depth"] < 64
    scene = orc.OracleScene((mesh.vertices, mesh.triangles))
    n_bad, n_all = _compare_single_axis(orc, scene, mesh, res, poses, intr)
    assert n_all >= 0.99 * n_pose * 128000
    assert n_bad / n_all <= 1e-5, (n_bad, n_all)
    if target_tris >= 10_000_000:
        engine.ctx.set_option("chunk_rays", 2 * 128000)                     # 3 chunks of 2 + 2 + 1 frames, double-buffered
        try:
            chunked = engine.simulate(poses, intr).numpy()
        finally:
            engine.ctx.set_option("chunk_rays", 1 << 26)
        for k in ("points", "incident", "prim_id", "label", "ray_idx", "frame_offset"):
            assert np.array_equal(chunked[k], res[k]), k
        host = engine.simulate_to_host(poses, intr, chunk_poses=2)
        assert np.array_equal(host["points"], res["points"]) and np.array_equal(host["incident"], res["incident"])
        assert np.array_equal(host["label"], res["label"]) and np.array_equal(host["frame_offset"], res["frame_offset"])
