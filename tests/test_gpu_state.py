"""GPU: state that outlives one call -- the grow-only scratch, the BVH cache, the context's single NN / collision index
slots and the gather regions.  Each test pins one way a second caller (or a second call) could silently change the
result of the first; the reference has no such state (it rebuilds its scene and fits its tree on every call,
raycast_engine_cpu.py:46-47, containers/s3dis_sim_scene.py:413-424)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pose(lrc, x=2.0, y=3.0, z=1.0, yaw=0.3):
    return lrc.poses_from_waypoints([lrc.Waypoint(x, y, z, yaw)])[0]


def test_one_frame_scan_sizes_scratch_by_the_frame(lrc):
    """A per-frame drop-in call (8 x 2000 rays) must not reserve the 2^26-ray default chunk (1.6 GB)."""
    ctx = lrc.Context(0)                       # fresh context: the scratch is grow-only
    try:
        mesh = lrc.synthetic.box_room(target_tris=6000, seed=3)
        ctx.set_mesh(mesh)
        intr = lrc.Indoor8LineLidarIntrinsics(max_range=10.0)
        build_scratch = ctx.stat("scratch_bytes")
        res = ctx.scan(_pose(lrc)[None], intr)
        assert res.num_points > 1000
        n = lrc.rays_per_frame(intr)
        grown = ctx.stat("scratch_bytes")
        assert grown <= max(build_scratch, 4 * 24 * n + (1 << 20)), (build_scratch, grown)
        # three frames in one call: still sized by the real chunk
        ctx.scan(np.stack([_pose(lrc, 2.0 + k) for k in range(3)]), intr)
        assert ctx.stat("scratch_bytes") <= max(build_scratch, 4 * 24 * 3 * n + (1 << 20))
    finally:
        ctx.close()


def test_in_place_mesh_edits_are_seen_by_the_cache(lrc):
    """One vertex or one label changed in place (anywhere in a large mesh) must rebuild, not reuse a stale BVH."""
    engine = lrc.RaycastEngineGPU()
    ctx = engine.ctx
    ctx.invalidate_mesh()
    mesh = lrc.synthetic.box_room(target_tris=60000, seed=9)
    lab = (np.arange(len(mesh.triangles)) % 13).astype(np.uint32)
    mesh = lrc.TriangleMesh(mesh.vertices, mesh.triangles, lab)
    lidar = lrc.create_lidar(lrc.Indoor8LineLidarIntrinsics(max_range=30.0), _pose(lrc))
    p0, _ = engine.lidar_intersect_mesh(lidar, mesh)
    lab0 = engine.last_scan.label.cpu().numpy().copy()
    prim0 = engine.last_scan.prim_id.cpu().numpy().copy()
    g0 = ctx.stat("mesh_generation")
    engine.lidar_intersect_mesh(lidar, mesh)
    assert ctx.stat("mesh_generation") == g0                    # unchanged content: BVH reused
    # relabel ONE triangle that the frame hits (an odd index: a strided sample of the buffer would skip it)
    hit = int(prim0[len(prim0) // 2])
    mesh.triangle_labels[hit] = 999
    engine.lidar_intersect_mesh(lidar, mesh)
    assert ctx.stat("mesh_generation") == g0 + 1
    lab1 = engine.last_scan.label.cpu().numpy()
    assert np.array_equal(lab1 == 999, prim0 == hit) and (prim0 == hit).any()
    assert np.array_equal(lab1[prim0 != hit], lab0[prim0 != hit])
    # move ONE vertex of that triangle by 5 cm along z: the points on it must change
    vi = int(mesh.triangles[hit][0])
    mesh.vertices[vi, 2] += 0.05
    p2, _ = engine.lidar_intersect_mesh(lidar, mesh)
    assert ctx.stat("mesh_generation") == g0 + 2
    assert p2.shape != p0.shape or not np.array_equal(p2, p0)
    # a fresh (vertices, triangles) tuple with the same content is the same mesh
    engine.lidar_intersect_mesh(lidar, (mesh.vertices.copy(), mesh.triangles.copy(), mesh.triangle_labels.copy()))
    assert ctx.stat("mesh_generation") == g0 + 2
    # cache_mesh=False is the reference's behaviour: rebuild on every call
    e2 = lrc.RaycastEngineGPU(cache_mesh=False)
    e2.lidar_intersect_mesh(lidar, mesh); e2.lidar_intersect_mesh(lidar, mesh)
    assert ctx.stat("mesh_generation") == g0 + 4


def test_pinned_mesh_skips_the_content_check_until_unpinned(lrc):
    engine = lrc.RaycastEngineGPU()
    ctx = engine.ctx
    mesh = lrc.synthetic.box_room(target_tris=8000, seed=4)
    other = lrc.synthetic.box_room(target_tris=8000, seed=5)
    lidar = lrc.create_lidar(lrc.Indoor8LineLidarIntrinsics(max_range=30.0), _pose(lrc))
    engine.set_mesh(mesh)                                       # build + pin
    g0 = ctx.stat("mesh_generation")
    a, _ = engine.lidar_intersect_mesh(lidar, mesh)
    assert ctx.stat("mesh_generation") == g0
    b, _ = engine.lidar_intersect_mesh(lidar, other)            # another object: un-pins, rebuilds
    assert ctx.stat("mesh_generation") == g0 + 1
    c, _ = engine.lidar_intersect_mesh(lidar, mesh)             # back: content differs from the resident BVH -> rebuild
    assert ctx.stat("mesh_generation") == g0 + 2
    assert np.array_equal(a, c) and not (a.shape == b.shape and np.array_equal(a, b))
    engine.set_mesh(mesh)
    engine.invalidate_mesh()
    engine.lidar_intersect_mesh(lidar, mesh)
    assert ctx.stat("mesh_generation") == g0 + 3


def test_two_label_transfers_do_not_share_an_index(lrc):
    """The context has ONE NN slot; a second LabelTransfer must not retarget the first one's queries."""
    ctx = lrc.get_context(0)
    rng = np.random.default_rng(11)
    pa = rng.uniform(0, 5, (400, 3)); sa = rng.integers(0, 13, 400)
    pb = rng.uniform(0, 5, (5000, 3)) + 1.0; sb = rng.integers(100, 113, 5000)
    q = rng.uniform(0, 5, (3000, 3)).astype(np.float32)

    def brute(ref):
        d = ((q.astype(np.float64)[:, None, :] - ref[None, :, :]) ** 2).sum(-1)
        return d.argmin(1)

    A = lrc.LabelTransfer(ctx, pa, semantic=sa)
    ia0 = A.query(q)["index"].cpu().numpy()
    B = lrc.LabelTransfer(ctx, pb, semantic=sb)            # re-targets the context's slot
    ra = A.query(q)
    ia, la = ra["index"].cpu().numpy(), ra["label"].cpu().numpy().view(np.uint32) & 0xFFFF
    assert np.array_equal(ia, ia0) and np.array_equal(ia, brute(pa)) and np.array_equal(la, sa[ia])
    rb = B.query(q)                                        # and B still answers with B's points after A took the slot back
    ib, lb = rb["index"].cpu().numpy(), rb["label"].cpu().numpy().view(np.uint32) & 0xFFFF
    assert np.array_equal(ib, brute(pb)) and np.array_equal(lb, sb[ib])


def test_two_planners_do_not_share_a_collision_index(lrc):
    from lrc_b200.trajectory.auto_trajectory_generator import AutoTrajectoryGenerator
    m1 = lrc.synthetic.box_room(target_tris=4000, seed=1)
    m2 = lrc.synthetic.planner_tight_room(seed=5)
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(0.2, 6.0, 500), rng.uniform(0.2, 5.0, 500), np.full(500, 0.5)], axis=1)
    g1, g2 = AutoTrajectoryGenerator(0.3), AutoTrajectoryGenerator(0.25)
    g1._index_mesh(m1)
    want1 = g1._query(pts, None)
    g2._index_mesh(m2)
    want2 = g2._query(pts, None)
    assert np.array_equal(g1._query(pts, None), want1)
    assert np.array_equal(g2._query(pts, None), want2)


def test_gathered_record_is_complete_and_regions_are_guarded(lrc):
    """world = 1 form of the exchange: the gather buffer receives xyz + label + offsets; the incident angles recomputed
    on arrival equal the scan's own bit for bit; a scan with more frames than the region holds is refused."""
    import torch
    from lrc_b200.distributed import PeerGather
    engine = lrc.RaycastEngineGPU()
    ctx = engine.ctx
    mesh = lrc.synthetic.box_room(target_tris=20000, seed=1)
    engine.set_mesh(mesh)
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    poses = lrc.poses_from_waypoints([lrc.Waypoint(1.5 + 0.5 * k, 3.2 + 0.1 * k, 1.0, 0.05 * k) for k in range(5)])
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=5)
    n = lrc.rays_per_frame(intr)
    pg = PeerGather(ctx, cap_per_rank=5 * n, frames_per_rank=5)
    try:
        pg.enable()
        ctx.set_option("gather_chunks", 2)
        for mode in (0, 1):                                   # exchange kernel: vector loads / stores, TMA bulk copies
            ctx.set_option("push_mode", mode)
            pg.buffer.zero_()
            ref = engine.simulate(poses, intr, noise=noise).numpy()
            pg.synchronize()
            got = pg.assemble_numpy(poses_all=poses)
            for k in ("frame_offset", "points", "label", "incident"):
                assert np.array_equal(got[k], ref[k]), (mode, k)
        assert got["incident"].dtype == np.float64 and len(got["incident"]) > 10000
        # compaction straight into this rank's own region (no self copy by the exchange kernel): same bits
        pg.buffer.zero_()
        direct = ctx.scan(poses, intr, noise, bufs=pg.local_out(5)).numpy()
        pg.synchronize()
        got = pg.assemble_numpy(poses_all=poses)
        for k in ("frame_offset", "points", "label", "incident"):
            assert np.array_equal(got[k], ref[k]), ("local_out", k)
        assert np.array_equal(direct["points"], ref["points"]) and np.array_equal(direct["incident"], ref["incident"])
        # compact wire format with a single rank: nothing to rebuild, the local cloud and its t | ray index arrays are written
        pg.enable(wire=True, poses_all=poses)
        pg.buffer.zero_()
        ctx.scan(poses, intr, noise, bufs=pg.local_out(5))
        pg.synchronize()
        got = pg.assemble_numpy(poses_all=poses)
        for k in ("frame_offset", "points", "label", "incident"):
            assert np.array_equal(got[k], ref[k]), ("wire", k)
        m = int(ref["frame_offset"][-1])
        ray = pg.buffer[pg.o_ray: pg.o_ray + 4 * m].view(torch.int32).cpu().numpy().view(np.uint32)
        assert np.array_equal(ray, ref["ray_idx"])
        with pytest.raises(RuntimeError, match="incident_deg"):
            ctx.scan(poses, intr, noise, bufs=pg.local_out(5, incident=True))
        pg.enable()
        more = lrc.poses_from_waypoints([lrc.Waypoint(1.5 + 0.4 * k, 3.2, 1.0, 0.0) for k in range(6)])
        with pytest.raises(RuntimeError, match="frame_capacity"):
            engine.simulate(more[:6], lrc.Indoor8LineLidarIntrinsics(max_range=5.0))
    finally:
        pg.disable()
        ctx.set_option("gather_chunks", 4)
        ctx.set_option("push_mode", 0)
        pg.close()


def test_per_frame_results_stay_valid_when_kept(lrc):
    """The per-waypoint call returns arrays backed by pooled page-locked buffers (8 per array kind).  A caller that keeps
    every frame (the reference's simulator does, s3dis_simulator.py:287) must still own independent, intact arrays: the
    pool falls back to ordinary copies once it is exhausted, and a dropped frame gives its buffer back."""
    engine = lrc.RaycastEngineGPU()
    mesh = lrc.synthetic.box_room(target_tris=20000, seed=1)
    engine.set_mesh(mesh)
    for intr in (lrc.Indoor8LineLidarIntrinsics(max_range=30.0), lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()):
        poses = lrc.poses_from_waypoints([lrc.Waypoint(1.5 + 0.3 * k, 3.2 + 0.05 * k, 1.0, 0.1 * k) for k in range(13)])
        dual = hasattr(intr, "num_vertical_lines")
        lidars = [lrc.create_lidar(intr, p) for p in poses]
        kept = [engine.lidar_intersect_mesh(l, mesh) for l in lidars]                   # 13 > 8 frames alive at once
        for k, (pts, inc) in enumerate(kept):
            nz = lidars[k].noise_config() if dual else None
            ref = engine.simulate(poses[k][None], intr, noise=nz).numpy()
            assert pts.dtype == np.float32 and inc.dtype == np.float64 and pts.flags.writeable
            assert np.array_equal(pts, ref["points"]) and np.array_equal(inc, ref["incident"]), k
        first = kept[0][0].copy()
        kept[5][0][:] = 0.0                                                              # frames do not alias each other
        assert np.array_equal(kept[0][0], first) and not np.array_equal(kept[4][0], kept[5][0])
        del kept
        again = engine.lidar_intersect_mesh(lidars[3], mesh)                             # buffers were returned and are reused
        ref = engine.simulate(poses[3][None], intr, noise=lidars[3].noise_config() if dual else None).numpy()
        assert np.array_equal(again[0], ref["points"]) and np.array_equal(again[1], ref["incident"])
