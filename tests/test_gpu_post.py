"""GPU: lrc_frame_statistics / lrc_pack_ply_records (csrc/post.cu) against the post-cast oracle and the fixtures the
reference's own writer produced.  Tolerances: the per-point float32 norm is bit-exact by construction; the reference
sums in float32 (range) / float64 (angles) with numpy's pairwise order, the GPU in float64 with a fixed tree --
range statistics agree to 2e-6 relative (float32 accumulation error of the reference), angle statistics to 1e-12."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _result(lrc, ctx, frames, labels=None, prims=None):
    import torch
    pts = np.concatenate([f[0] for f in frames]) if frames else np.zeros((0, 3), np.float32)
    inc = np.concatenate([f[1] for f in frames]) if frames else np.zeros(0)
    off = np.concatenate([[0], np.cumsum([len(f[0]) for f in frames])]).astype(np.int64)
    dev = ctx.device
    e32 = torch.empty(0, dtype=torch.int32, device=dev)

    def i32(a):
        return e32 if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).to(dev)
    return lrc.ScanResult(points=torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).to(dev).reshape(-1, 3),
                          incident=torch.from_numpy(np.ascontiguousarray(inc, dtype=np.float64)).to(dev),
                          prim_id=i32(prims), label=i32(labels), ray_idx=e32,
                          frame_offset=torch.from_numpy(off).to(dev), frame_offset_host=off)


def test_frame_statistics_vs_reference_fixture(lrc, golden):
    ctx = lrc.get_context(0)
    g = golden("post_stats.npz")
    frames = [(g[f"frame{i}/points"], g[f"frame{i}/incident"]) for i in range(3)]
    res = _result(lrc, ctx, frames)
    qs = lrc.scan_quality(ctx, res, int(g["total_points_per_scan"]), float(g["room_volume"]))
    for i, q in enumerate(qs):
        want = g[f"frame{i}/quality"]          # coverage, n, inc mean, inc std, density, range mean, range std
        assert q.num_points == int(want[1]) and q.coverage_ratio == want[0] and q.scan_density == want[4]
        assert q.incident_angle_mean == pytest.approx(want[2], rel=1e-12, abs=1e-12)
        assert q.incident_angle_std == pytest.approx(want[3], rel=1e-10, abs=1e-12)
        assert q.range_mean == pytest.approx(want[5], rel=2e-6, abs=1e-12)
        assert q.range_std == pytest.approx(want[6], rel=2e-5, abs=1e-12)
    s = lrc.simulation_stats(qs, 2.5)
    want = g["stats"]
    assert s.total_frames == 3 and s.total_points == int(want[1]) and s.frames_per_second == want[7]
    assert s.average_incident_angle == pytest.approx(want[4], rel=1e-12) and s.average_range == pytest.approx(want[5], rel=2e-6)


def test_frame_statistics_on_a_real_scan_and_float64_oracle(lrc):
    """Statistics of an actual trajectory scan vs numpy in float64 on the same points (tight), incl. determinism."""
    from oracle import post_oracle as po
    eng = lrc.RaycastEngineGPU(device=0)
    mesh = lrc.synthetic.box_room(target_tris=6000, seed=3)
    poses = lrc.poses_from_waypoints([lrc.Waypoint(3.0 + 0.3 * k, 2.7, 1.0, 0.1 * k) for k in range(5)])
    intr = lrc.Indoor8LineLidarIntrinsics(max_range=4.0, horizontal_res=700)
    res = eng.simulate(poses, intr, mesh)
    st = lrc.frame_statistics(eng.ctx, res)
    st2 = lrc.frame_statistics(eng.ctx, res)
    assert st.tobytes() == st2.tobytes()                                    # fixed summation order
    r = res.numpy()
    for p in range(5):
        a, b = r["frame_offset"][p], r["frame_offset"][p + 1]
        pts, inc = r["points"][a:b], r["incident"][a:b]
        q = po.scan_quality(pts, inc, 8 * 700, 240.0)
        rng = np.linalg.norm(pts, axis=1).astype(np.float64)                # float32 norms, float64 statistics
        assert st["num_points"][p] == b - a == q["num_points"]
        assert st["incident_mean"][p] == pytest.approx(inc.mean(), rel=1e-13)
        assert st["incident_std"][p] == pytest.approx(inc.std(), rel=1e-10)
        assert st["range_mean"][p] == pytest.approx(rng.mean(), rel=1e-13)
        assert st["range_std"][p] == pytest.approx(rng.std(), rel=1e-10)
        assert st["range_mean"][p] == pytest.approx(float(q["range_mean"]), rel=2e-6)   # the reference's float32 mean


def test_ply_records_match_reference_writer_bytes(lrc, golden, tmp_path):
    import torch
    ctx = lrc.get_context(0)
    g = golden("post_stats.npz")
    pts, colors, sem, ins = g["ply/points"], g["ply/colors"], g["ply/sem"], g["ply/ins"]
    n = len(pts)
    # colours arrive through the per-triangle table: give every point its own "triangle"
    prims = np.arange(n, dtype=np.uint32)[::-1].copy()
    tri_rgb = np.zeros(n, np.uint32)
    tri_rgb[prims] = lrc.post.pack_rgb(colors)
    res = _result(lrc, ctx, [(pts, np.zeros(n))], labels=lrc.pack_labels(sem, ins), prims=prims)
    path = tmp_path / "out.ply"
    nbytes = lrc.write_labeled_ply(ctx, path, res, tri_rgb=tri_rgb)
    want = open(os.path.join(GOLDEN, "post_labeled.ply"), "rb").read()
    got = open(path, "rb").read()
    assert nbytes == len(want) and got == want                              # byte-identical to the reference's file


def test_ply_default_colour_empty_and_large(lrc, tmp_path):
    from oracle import post_oracle as po
    ctx = lrc.get_context(0)
    rng = np.random.default_rng(5)
    # default grey + default labels (reference :584-594), sizes around the 256-point tile
    for n in (0, 1, 255, 256, 513):
        pts = rng.standard_normal((n, 3)).astype(np.float32)
        res = _result(lrc, ctx, [(pts, np.zeros(n))])
        path = tmp_path / f"d{n}.ply"
        lrc.write_labeled_ply(ctx, path, res)
        want = po.labeled_ply_bytes(pts, np.full((n, 3), 127, np.uint8), np.zeros(n, np.uint16), np.zeros(n, np.uint16))
        assert open(path, "rb").read() == want
    # 2M points: vectorised structured-dtype packing of the same fields as the checker
    n = 2_000_003
    pts = rng.standard_normal((n, 3)).astype(np.float32)
    lab = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    res = _result(lrc, ctx, [(pts, np.zeros(n))], labels=lab)
    path = tmp_path / "big.ply"
    lrc.write_labeled_ply(ctx, path, res, default_rgb=0x030201)
    rec = lrc.read_labeled_ply(path)
    assert np.array_equal(rec["points"].view(np.uint32), pts.view(np.uint32))
    assert np.array_equal(rec["semantic_labels"], (lab & 0xFFFF).astype(np.uint16))
    assert np.array_equal(rec["instance_labels"], (lab >> 16).astype(np.uint16))
    assert (rec["colors"] == np.array([1, 2, 3], np.uint8)).all()
