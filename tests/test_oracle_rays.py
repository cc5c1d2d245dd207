"""Oracle ray tables vs the golden vectors captured from the LIVE reference `lidar` package
(tests/golden/make_golden.py; reference lidar/indoor_lidar.py:56-131,224-296)."""
import hashlib

import numpy as np
import pytest

from conftest import ulp_diff

SINGLE = ["8line", "32line", "64line", "lowcost", "custom"]
STRIDE = 37


@pytest.mark.parametrize("preset", SINGLE)
@pytest.mark.parametrize("pose_name", ["identity", "posed"])
def test_single_axis_table_matches_reference(orc, golden, golden_poses, preset, pose_name):
    g = golden("rays_single_axis.npz")
    vd = list(g[f"{preset}/vertical_degrees"])
    W = int(g[f"{preset}/W"])
    rays = orc.gen_rays_single_axis(golden_poses[pose_name], vd, W)
    assert rays.dtype == np.float32 and rays.shape == (int(g[f"{preset}/{pose_name}/n"]), 6)
    sample = g[f"{preset}/{pose_name}/sample"]
    d = ulp_diff(rays[::STRIDE], sample)
    assert d.max() <= 1                                  # float64 libm differences may flip a float32 rounding
    assert (d > 0).mean() <= 1e-4
    np.testing.assert_allclose(rays.astype(np.float64).sum(0), g[f"{preset}/{pose_name}/sum"], rtol=0, atol=1e-3)
    # on this toolchain the table is bit-identical to the reference's; keep that visible without making the
    # suite depend on libm's last bit
    same = hashlib.sha256(rays.tobytes()).digest() == g[f"{preset}/{pose_name}/sha256"].tobytes()
    if not same:
        pytest.xfail("float32 table differs from the reference in the last bit somewhere (libm)")


def test_ray_order_is_line_major(orc):
    """index = j*W + i, azimuth beta = -(i - W/2)/W*2pi: i = W/2 looks along +x (reference :108-113)."""
    W = 8
    rays = orc.gen_rays_single_axis(np.eye(4), [10.0, -10.0], W)
    d = rays[:, 3:]
    assert d[W // 2, 0] == pytest.approx(np.cos(np.deg2rad(10.0)), abs=1e-7) and abs(d[W // 2, 1]) < 1e-7
    assert d[0, 0] == pytest.approx(-np.cos(np.deg2rad(10.0)), abs=1e-7)     # i = 0 -> beta = +pi
    assert np.all(d[:W, 2] > 0) and np.all(d[W:, 2] < 0)
    assert d[W // 2 - 1, 1] > 0                                               # azimuth decreases with i
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=2e-7)


def test_single_axis_empty_table_defaults_to_one_level_line(orc):
    rays = orc.gen_rays_single_axis(np.eye(4), [], 5)                        # reference :104-106
    assert rays.shape == (5, 6) and np.all(rays[:, 5] == 0)


@pytest.mark.parametrize("key", ["1x16/identity", "1x16/posed", "4x50/identity", "4x50/posed", "8x360/identity", "8x360/posed"])
def test_uniform_fov_table_matches_reference(orc, golden, golden_poses, key):
    g = golden("rays_uniform.npz")
    hw, pose_name = key.split("/")
    H, W = map(int, hw.split("x"))
    rays = orc.gen_rays_uniform(golden_poses[pose_name], 15.0, 20.0, H, W)
    assert ulp_diff(rays, g[key + "/rays"]).max() <= 1


@pytest.mark.parametrize("pose_name", ["identity", "posed"])
def test_dual_axis_noise_free_table_matches_reference(orc, lrc, golden, golden_poses, pose_name):
    g = golden("rays_dual_axis.npz")
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    rays, keep = orc.gen_rays_dual_axis(golden_poses[pose_name], orc.dual_params(intr, 0.0, 0.0))
    assert rays.shape == (64000, 6) and keep.all()
    d = ulp_diff(rays[::STRIDE], g[f"blk2go/{pose_name}/sample"])
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4
    np.testing.assert_allclose(rays.astype(np.float64).sum(0), g[f"blk2go/{pose_name}/sum"], atol=1e-3)


def test_dual_axis_small_irregular_sensor(orc, lrc, golden, golden_poses):
    g = golden("rays_dual_axis.npz")
    small = lrc.DualAxisLidarIntrinsics(point_rate=1000, scan_duration=0.5, num_vertical_lines=7, swing_frequency=3.0,
                                        swing_amplitude=0.3, angle_noise_std=0.0, dropout_probability=0.0)
    rays, _ = orc.gen_rays_dual_axis(golden_poses["posed"], orc.dual_params(small))
    assert rays.shape == g["small/rays"].shape == (7 * (500 // 7), 6)
    assert ulp_diff(rays, g["small/rays"]).max() <= 1


def test_dual_axis_noise_and_dropout_statistics(orc, lrc):
    """With noise on, results are statistically (not bitwise) comparable with the reference's numpy stream."""
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    clean, _ = orc.gen_rays_dual_axis(np.eye(4), orc.dual_params(intr, 0.0, 0.0))
    noisy, keep = orc.gen_rays_dual_axis(np.eye(4), orc.dual_params(intr), seed=7, pose_idx=3, compact=False)
    assert abs(keep.mean() - 0.98) < 0.003                                   # dropout_probability = 0.02
    ang = np.arccos(np.clip((clean[:, 3:] * noisy[:, 3:]).sum(1), -1, 1))
    # phi and theta each get N(0, 1e-3): angular deviation ~ Rayleigh-like with scale ~1e-3
    assert 0.8e-3 < np.sqrt((ang ** 2).mean() / 2) < 1.2e-3
    again, keep2 = orc.gen_rays_dual_axis(np.eye(4), orc.dual_params(intr), seed=7, pose_idx=3, compact=False)
    assert np.array_equal(noisy, again) and np.array_equal(keep, keep2)      # counter-based: reproducible
    other, _ = orc.gen_rays_dual_axis(np.eye(4), orc.dual_params(intr), seed=7, pose_idx=4, compact=False)
    assert not np.array_equal(noisy, other)


def test_philox_known_answer(orc):
    """Philox4x32-10 test vectors from the Random123 distribution (kat_vectors)."""
    import ctypes
    L = orc.lib()
    out = np.zeros(4, np.uint32)
    # counter = (ray, pose_lo, pose_hi, stream), key = seed
    L.orc_philox(0, 0, 0, 0, out.ctypes.data_as(ctypes.c_void_p))
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.orc_philox(0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, out.ctypes.data_as(ctypes.c_void_p))
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
