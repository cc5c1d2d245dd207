"""GPU: out-of-bounds guards.  compute-sanitizer is closed on this GPU pool, so the kernels that scatter through
computed offsets are checked the other way round: every output lives inside a larger canary-filled allocation, the
capacity handed to the library is exact, and the canaries on both sides must survive."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CANARY = 0x5A


def _guarded(torch, dev, shape, dtype, pad=4096):
    n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
    raw = torch.full((pad + n + pad,), CANARY, dtype=torch.uint8, device=dev)
    view = raw[pad:pad + n].view(dtype).view(*shape)
    return raw, view, pad, n


def _intact(raw, pad, n):
    return bool((raw[:pad] == CANARY).all().item()) and bool((raw[pad + n:] == CANARY).all().item())


def test_scan_outputs_stay_inside_their_buffers(engine, lrc):
    import torch
    ctx = engine.ctx
    dev = ctx.device
    mesh = lrc.synthetic.box_room(target_tris=6000, seed=3)
    engine.set_mesh(mesh)
    for intr, noise in ((lrc.Indoor8LineLidarIntrinsics(max_range=4.0, horizontal_res=999), None),
                        (lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis(), lrc.NoiseConfig(0.001, 0.3, 0.0, 5, 0))):
        poses = lrc.poses_from_waypoints([lrc.Waypoint(2.0 + k, 3.0, 1.0, 0.2 * k) for k in range(3)])
        P, N = 3, lrc.rays_per_frame(intr)
        cap = P * N
        g = {k: _guarded(torch, dev, shp, dt) for k, (shp, dt) in {
            "xyz": ((cap, 3), torch.float32), "incident": ((cap,), torch.float64), "prim": ((cap,), torch.int32),
            "label": ((cap,), torch.int32), "ray": ((cap,), torch.int32), "off": ((P + 1,), torch.int64)}.items()}
        bufs = {k: v[1] for k, v in g.items()}
        for chunk in (1 << 26, N):                                     # one launch, then one frame per chunk (pipelined)
            ctx.set_option("chunk_rays", chunk)
            ctx.scan_enqueue(torch.from_numpy(poses.reshape(-1, 16)).to(dev), intr, noise, bufs)
            torch.cuda.synchronize()
            for k, (raw, _, pad, n) in g.items():
                assert _intact(raw, pad, n), (type(intr).__name__, chunk, k)
        ctx.set_option("chunk_rays", 1 << 26)
        m = int(bufs["off"][-1].item())
        assert 0 < m <= cap


def test_post_and_nn_outputs_stay_inside_their_buffers(lrc):
    import ctypes as C
    import torch
    from lrc_b200 import _native as nat
    ctx = lrc.get_context(0)
    dev = ctx.device
    rng = np.random.default_rng(0)
    for m in (1, 255, 257, 5000):
        pts = torch.from_numpy(rng.standard_normal((m, 3)).astype(np.float32)).to(dev)
        lab = torch.from_numpy(rng.integers(0, 2**31, m).astype(np.int32)).to(dev)
        raw, rec, pad, n = _guarded(torch, dev, (19 * m,), torch.uint8, pad=4096)
        nat.check(ctx._h, ctx._lib.lrc_pack_ply_records(ctx._h, C.c_void_p(pts.data_ptr()), C.c_void_p(lab.data_ptr()), None, None,
                                                        0x7F7F7F, m, C.c_void_p(rec.data_ptr()), None))
        torch.cuda.synchronize()
        assert _intact(raw, pad, n), m
        # 1-NN: index / distance / two attribute outputs
        lt = lrc.LabelTransfer(ctx, rng.standard_normal((300, 3)), semantic=rng.integers(0, 13, 300), colors=rng.random((300, 3)))
        outs = [_guarded(torch, dev, (m,), dt) for dt in (torch.int32, torch.float64, torch.int32, torch.int32)]
        nat.check(ctx._h, ctx._lib.lrc_nn_query(ctx._h, C.c_void_p(pts.data_ptr()), m, C.c_void_p(outs[0][1].data_ptr()),
                                                C.c_void_p(outs[1][1].data_ptr()), C.c_void_p(lt._lab.data_ptr()), C.c_void_p(outs[2][1].data_ptr()),
                                                C.c_void_p(lt._rgb.data_ptr()), C.c_void_p(outs[3][1].data_ptr()), None))
        torch.cuda.synchronize()
        for raw, _, pad, n in outs:
            assert _intact(raw, pad, n), m
        assert int(outs[0][1].min()) >= 0 and int(outs[0][1].max()) < 300
