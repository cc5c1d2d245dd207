"""GPU: randomised property tests (hypothesis) -- arbitrary triangle soups and ray bundles, LBVH traversal against the
oracle's exhaustive float32 search (same intersection spec => identical bits) and against the GPU's own exhaustive
kernel; ragged sizes around the 32 / 128 / 4096 tiling boundaries."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu
MISS = 0xFFFFFFFF


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(seed=st.integers(0, 2**31 - 1), n_tri=st.sampled_from([1, 2, 3, 31, 33, 127, 129, 1000, 4097]),
       n_ray=st.sampled_from([0, 1, 31, 32, 33, 127, 128, 129, 1000]), scale=st.sampled_from([1e-3, 1.0, 250.0]),
       flat=st.booleans(), node_format=st.sampled_from([0, 1, 2]), persistent=st.sampled_from([0, 1, 2]), rpt=st.sampled_from([1, 2, 4]), compact=st.booleans(), leaf_size=st.sampled_from([1, 3, 4, 8]),
       quality=st.sampled_from([0, 1]), radius=st.sampled_from([1, 5, 16, 32]), variant=st.sampled_from([1, 65]))
def test_random_soups_bvh_equals_exhaustive(engine, lrc, orc, seed, n_tri, n_ray, scale, flat, node_format, leaf_size, quality,
                                            radius, variant, persistent, rpt, compact):
    engine.ctx.set_option("node_format", node_format)          # 64 B float boxes / 32 B 16-bit boxes: identical results
    engine.ctx.set_option("leaf_size", leaf_size)              # 1..8 triangles per leaf: identical results
    engine.ctx.set_option("build_quality", quality)            # LBVH / PLOC: identical results
    engine.ctx.set_option("ploc_radius", radius)
    engine.ctx.set_option("variant", variant)                  # plain stack / stack entries culled at pop time
    engine.ctx.set_option("persistent", persistent)            # one block per 128 rays / persistent warps over 32-ray tiles
    engine.ctx.set_option("compact_nodes", int(compact))        # dead node records kept / squeezed out: identical results
    engine.ctx.set_option("rays_per_thread", rpt)              # 1 / 2 / 4 adjacent rays per thread (format 2 only): identical results
    engine.ctx.invalidate_mesh()
    rng = np.random.default_rng(seed)
    centre = rng.standard_normal((n_tri, 1, 3)) * scale
    verts = (centre + rng.standard_normal((n_tri, 3, 3)) * scale * rng.choice([0.01, 0.3])).reshape(-1, 3)
    if flat:
        verts[:, 2] = 0.25 * scale                                   # coplanar soup: many equal-t candidates, ties by id
    verts = verts.astype(np.float32).astype(np.float64)
    tris = np.arange(3 * n_tri, dtype=np.int32).reshape(-1, 3)
    if n_tri > 3:
        tris[1] = tris[0]                                           # duplicate triangle: exact tie -> smaller id wins
        tris[2, 2] = tris[2, 1]                                     # degenerate triangle: never hit
    o = rng.standard_normal((n_ray, 3)) * scale * 2
    d = rng.standard_normal((n_ray, 3)) * rng.choice([1.0, 1e-3, 40.0])
    if n_ray > 2:
        d[0] = (0, 0, -1.0); d[1, 0] = 0.0                          # axis-parallel components (1/0 handling in the slab test)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    mesh = lrc.TriangleMesh(verts, tris)
    t, pid = engine.cast_rays(rays, mesh)
    tb, pb = (x.cpu().numpy() for x in engine.ctx.cast_rays(rays, bruteforce=True))
    assert np.array_equal(t.view(np.uint32), tb.view(np.uint32)) and np.array_equal(pid, pb.view(np.uint32))
    to, po = orc.cast_rays_brute((verts, tris), rays)
    assert np.array_equal(t.view(np.uint32), np.asarray(to, np.float32).view(np.uint32)) and np.array_equal(pid, po)
    if n_tri > 3 and n_ray:
        assert not (pid == 1).any() and not (pid == 2).any()
    pts = engine.rays_intersect_mesh(rays, mesh)                # the scan-type kernel (packets when rays_per_thread > 1)
    assert len(pts) == int((pid != MISS).sum())
    scan = engine.last_scan.numpy()
    hit = pid != MISS
    assert np.array_equal(scan["prim_id"], pid[hit]) and np.array_equal(scan["ray_idx"], np.nonzero(hit)[0].astype(np.uint32))
    engine.ctx.set_option("node_format", 2)                     # back to the defaults
    engine.ctx.set_option("leaf_size", 2)
    engine.ctx.set_option("build_quality", 0)
    engine.ctx.set_option("ploc_radius", 16)
    engine.ctx.set_option("compact_nodes", 0)
    engine.ctx.set_option("variant", 65)
    engine.ctx.set_option("persistent", 0)
    engine.ctx.set_option("rays_per_thread", 1)
    engine.ctx.invalidate_mesh()
