"""GPU: error behaviour of the C ABI (include/lrc.h): every failure is a negative status + message, nothing is swallowed
(the reference's caller hides engine errors as empty frames, s3dis_simulator.py:271-273 -- the drop-in must not)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fresh(lrc):
    from lrc_b200.core import Context
    return Context(0)


def test_status_codes_and_messages(lrc):
    import torch
    from lrc_b200 import _native as nat
    ctx = _fresh(lrc)
    lib, h = ctx._lib, ctx._h
    dev = ctx.device
    rays = torch.zeros((4, 6), dtype=torch.float32, device=dev)
    t = torch.zeros(4, dtype=torch.float32, device=dev)
    pid = torch.zeros(4, dtype=torch.int32, device=dev)
    # cast before any mesh
    rc = lib.lrc_cast_rays(h, C.c_void_p(rays.data_ptr()), 4, C.c_void_p(t.data_ptr()), C.c_void_p(pid.data_ptr()), None)
    assert rc == -3 and b"lrc_set_mesh" in lib.lrc_last_error(h)
    with pytest.raises(nat.LrcError) as e:
        ctx.bvh_info()
    assert e.value.code == -3
    # triangle index out of range, NaN vertex, negative sizes
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    with pytest.raises(nat.LrcError) as e:
        ctx.set_mesh_arrays(v, np.array([[0, 1, 3]], np.int32))
    assert e.value.code == -1 and "outside" in str(e.value)
    bad = v.copy(); bad[1, 1] = np.nan
    with pytest.raises(nat.LrcError) as e:
        ctx.set_mesh_arrays(bad, np.array([[0, 1, 2]], np.int32))
    assert e.value.code == -1 and "non-finite" in str(e.value)
    assert lib.lrc_set_mesh(h, None, -1, None, 0, None, None) == -1
    # a good mesh, then: output capacity too small, missing required arrays, unknown option, bad option value
    ctx.set_mesh_arrays(v, np.array([[0, 1, 2]], np.int32))
    assert ctx.bvh_info()["num_tris"] == 1
    intr = lrc.Indoor8LineLidarIntrinsics.create_standard_8line()
    poses = torch.from_numpy(np.eye(4).reshape(1, 16)).to(dev)
    bufs, _ = ctx._alloc_out(100, 1)                                   # 16000 rays need 16000 slots
    with pytest.raises(nat.LrcError) as e:
        ctx.scan_enqueue(poses, intr, None, bufs)
    assert e.value.code == -4 and "capacity" in str(e.value)
    out = nat.Out(None, None, None, None, None, None, 16000)
    d = lrc.core.single_axis_desc(intr)
    assert lib.lrc_scan_single_axis(h, C.c_void_p(poses.data_ptr()), 1, C.byref(d), None, C.byref(out), None) == -1
    with pytest.raises(nat.LrcError):
        ctx.set_option("no_such_knob", 1)
    with pytest.raises(nat.LrcError):
        ctx.set_option("variant", 4)
    with pytest.raises(nat.LrcError):
        ctx.set_option("block", 96)
    for key, bad in (("tune", 3), ("push_tile", 1000), ("push_mode", 2), ("scan_chunks", 0), ("gather_taper", 0)):
        with pytest.raises(nat.LrcError):
            ctx.set_option(key, bad)
    # state inspection: known keys answer, unknown ones fail; incident angles / wire format reject bad arguments
    assert ctx.stat("node_format") == 2 and ctx.stat("num_sms") > 0 and ctx.stat("mesh_generation") >= 1
    with pytest.raises(nat.LrcError):
        ctx.stat("no_such_stat")
    assert lib.lrc_incident_angles(h, None, None, 1, None, 5, None, None) == -1
    assert lib.lrc_incident_angles(h, None, None, 0, None, 0, None, None) == 0
    gw = nat.GatherWire()
    gw.enabled = 1
    assert lib.lrc_set_gather_wire(h, C.byref(gw)) == -1 and b"lrc_set_gather first" in lib.lrc_last_error(h)
    assert lib.lrc_set_gather_wire(h, None) == 0
    # the context is still usable after all those failures
    res = ctx.scan(np.eye(4)[None], intr)
    assert res.num_frames == 1
    ctx.close()


def test_no_context_and_bad_device(lrc):
    from lrc_b200 import _native as nat
    lib = nat.load()
    h = C.c_void_p()
    assert lib.lrc_create(9999, C.byref(h)) == -1 and not h.value
    assert b"device index" in lib.lrc_last_error(None)
    assert lib.lrc_create(0, None) == -1
    assert lib.lrc_set_counting(None, 1) == -1 and lib.lrc_launch_count(None) == 0
    lib.lrc_destroy(None)                                               # a no-op, must not crash


def test_planner_and_nn_preconditions(lrc):
    import torch
    from lrc_b200 import _native as nat
    ctx = _fresh(lrc)
    lib, h = ctx._lib, ctx._h
    q = torch.zeros((2, 3), dtype=torch.float64, device=ctx.device)
    st = torch.zeros(2, dtype=torch.uint8, device=ctx.device)
    assert lib.lrc_collision_query(h, C.c_void_p(q.data_ptr()), 2, 0.3, None, C.c_void_p(st.data_ptr()), None) == -3
    qf = torch.zeros((2, 3), dtype=torch.float32, device=ctx.device)
    idx = torch.zeros(2, dtype=torch.int32, device=ctx.device)
    assert lib.lrc_nn_query(h, C.c_void_p(qf.data_ptr()), 2, C.c_void_p(idx.data_ptr()), None, None, None, None, None, None) == -3
    assert lib.lrc_collision_index_build(h, None, 5, 0.6, None) == -1
    assert lib.lrc_collision_index_build(h, C.c_void_p(q.data_ptr()), 2, 0.0, None) == -1
    # tri_rgb without prim_id, misaligned PLY output
    out = torch.zeros(19 * 2 + 16, dtype=torch.uint8, device=ctx.device)
    assert lib.lrc_pack_ply_records(h, C.c_void_p(qf.data_ptr()), None, None, C.c_void_p(idx.data_ptr()), 0, 2, C.c_void_p(out.data_ptr()), None) == -1
    assert lib.lrc_pack_ply_records(h, C.c_void_p(qf.data_ptr()), None, None, None, 0, 2, C.c_void_p(out.data_ptr() + 1), None) == -1
    ctx.close()
