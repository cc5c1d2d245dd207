"""GPU: exact 1-NN label / colour transfer (csrc/nn.cu) against scikit-learn's ball tree -- the reference's own
third-party dependency for this step (containers/s3dis_sim_scene.py:413-424), which IS installed here -- and against a
float64 brute-force argmin.  Indices are bit-exact (the synthetic data has no exact distance ties)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _annotated(rng, n):
    """S3DIS-like annotated points: samples on the faces of a room and of a few boxes, millimetre-rounded like the
    dataset's text files, with labels and colours."""
    pts = []
    for (lo, hi) in (((0, 0, 0), (10, 8, 3)), ((1, 1, 0), (2.6, 1.8, 0.75)), ((4, 7.5, 0), (6, 7.9, 2))):
        lo, hi = np.array(lo, float), np.array(hi, float)
        m = n // 3
        p = lo + rng.random((m, 3)) * (hi - lo)
        ax = rng.integers(0, 3, m)
        side = rng.integers(0, 2, m)
        p[np.arange(m), ax] = np.where(side == 0, lo[ax], hi[ax])
        pts.append(p)
    pts = np.round(np.concatenate(pts), 3)
    pts = np.unique(pts, axis=0)                                   # no duplicate points: no exact ties by construction
    rng.shuffle(pts)
    return pts, rng.random((len(pts), 3)), rng.integers(0, 13, len(pts)), rng.integers(0, 60000, len(pts))


def test_nn_transfer_matches_sklearn_ball_tree(lrc):
    from sklearn.neighbors import NearestNeighbors
    ctx = lrc.get_context(0)
    rng = np.random.default_rng(0)
    ref, colors, sem, ins = _annotated(rng, 60000)
    # queries: real hit points of a scan of the same room (float32), plus outliers far outside the annotated cloud
    eng = lrc.RaycastEngineGPU(device=0)
    mesh = lrc.synthetic.box_room(target_tris=6000, seed=3)
    res = eng.simulate(lrc.poses_from_waypoints([lrc.Waypoint(3.1, 2.7, 1.0, 0.3), lrc.Waypoint(6.0, 4.0, 1.0, 1.0)]),
                       lrc.Indoor8LineLidarIntrinsics.create_standard_8line(), mesh)
    q = np.concatenate([res.points.cpu().numpy(), (rng.standard_normal((500, 3)) * 30).astype(np.float32)])
    nbrs = NearestNeighbors(n_neighbors=1, algorithm="ball_tree").fit(ref)
    want_d, want_i = nbrs.kneighbors(q)
    lt = lrc.LabelTransfer(ctx, ref, colors=colors, semantic=sem, instance=ins)
    got = lt.query(q, want_distance=True)
    idx = got["index"].cpu().numpy()
    assert np.array_equal(idx, want_i.ravel())                                    # bit-exact indices
    assert np.allclose(got["distance"].cpu().numpy(), want_d.ravel(), rtol=1e-12, atol=0)
    # what the reference derives from the indices (:419-421, :483)
    lab = got["label"].cpu().numpy().view(np.uint32)
    assert np.array_equal(lab & 0xFFFF, sem[want_i.ravel()]) and np.array_equal(lab >> 16, ins[want_i.ravel()])
    rgb = got["rgb"].cpu().numpy().view(np.uint32)
    c255 = (colors[want_i.ravel()] * 255).astype(np.uint8)
    assert np.array_equal(np.stack([rgb & 255, (rgb >> 8) & 255, (rgb >> 16) & 255], 1).astype(np.uint8), c255)
    # relabel + PLY: the records carry the neighbour's colour and labels
    rl = lt.relabel(res)
    rec = lrc.post.pack_ply_records(ctx, rl, lt.rgb_table).cpu().numpy().view(lrc.post.PLY_DTYPE)
    m = res.num_points
    assert np.array_equal(rec["sem"], sem[want_i.ravel()[:m]].astype(np.uint16)) and np.array_equal(rec["red"], c255[:m, 0])


def test_nn_brute_force_ties_and_edge_cases(lrc):
    import torch
    ctx = lrc.get_context(0)
    rng = np.random.default_rng(1)
    ref = rng.standard_normal((5000, 3)) * np.array([5.0, 0.01, 2.0])             # very anisotropic cloud
    ref[100] = ref[7]                                                             # duplicate point: an exact tie
    q = np.concatenate([ref[:300].astype(np.float32), (rng.standard_normal((700, 3)) * 6).astype(np.float32)])
    for cell in (0.0, 0.05, 3.0):
        lt = lrc.LabelTransfer(ctx, ref, cell=cell)
        idx = lt.query(q)["index"].cpu().numpy()
        d = ((q.astype(np.float64)[:, None, :] - ref[None, :, :]) ** 2)
        rd = (d[:, :, 0] + d[:, :, 1]) + d[:, :, 2]
        assert np.array_equal(idx, rd.argmin(axis=1))                             # argmin returns the first minimum
    assert idx[7] == 7 and idx[100] == 7                                          # tie -> smaller index
    # empty index, empty query, non-finite query, non-finite annotated point
    assert lrc.LabelTransfer(ctx, np.zeros((0, 3))).query(q[:5])["index"].cpu().numpy().tolist() == [-1] * 5
    assert lt.query(np.zeros((0, 3), np.float32))["index"].numel() == 0
    bad = q[:3].copy(); bad[1, 2] = np.nan
    assert lt.query(bad)["index"].cpu().numpy()[1] == -1
    ref2 = ref.copy(); ref2[0] = np.inf
    assert (lrc.LabelTransfer(ctx, ref2).query(q)["index"].cpu().numpy() != 0).all()


def test_nn_at_scale_properties(lrc):
    """2M annotated points x 4M queries: every annotated point is its own neighbour; the distance returned for random
    queries is not beaten by any of 64 random annotated points (necessary condition), and a sample is checked exactly."""
    import torch
    ctx = lrc.get_context(0)
    g = torch.Generator(device="cpu").manual_seed(0)
    ref = (torch.rand((2_000_000, 3), generator=g, dtype=torch.float64) * torch.tensor([20.0, 15.0, 3.0], dtype=torch.float64)).numpy()
    lt = lrc.LabelTransfer(ctx, ref)
    self_idx = lt.query(ref.astype(np.float32), want_distance=True)
    # float32 rounding of the query moves it by < 2e-6 m; the neighbour is still the point itself unless another point is that close
    same = (self_idx["index"].cpu().numpy() == np.arange(len(ref))).mean()
    assert same > 0.9999
    q = (torch.rand((4_000_000, 3), generator=g) * torch.tensor([20.0, 15.0, 3.0])).numpy()
    out = lt.query(q, want_distance=True)
    idx, dist = out["index"].cpu().numpy(), out["distance"].cpu().numpy()
    assert (idx >= 0).all()
    rng = np.random.default_rng(2)
    rnd = ref[rng.integers(0, len(ref), 64)]
    for k in range(0, 64, 8):
        dd = np.sqrt(((q[::1000, None, :].astype(np.float64) - rnd[None, k:k + 8, :]) ** 2).sum(-1)).min(1)
        assert (dist[::1000] <= dd + 1e-12).all()
    sel = rng.integers(0, len(q), 200)
    d = ((q[sel].astype(np.float64)[:, None, :] - ref[None, :, :]) ** 2)
    assert np.array_equal(idx[sel], ((d[:, :, 0] + d[:, :, 1]) + d[:, :, 2]).argmin(1))
