#!/usr/bin/env python
"""Where the time of one per-waypoint `lidar_intersect_mesh` call goes (32-line sensor, 1M-triangle office, mesh pinned):
sensor construction, the engine call as the user sees it (results land in pooled page-locked buffers that become the numpy
arrays), the library call alone (kernels + D2H + the one synchronisation), what host copies of the two arrays WOULD cost
(the fallback when the caller keeps more than 8 frames alive), and the kernels alone."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lrc_b200 as lrc  # noqa: E402
from lrc_b200 import core  # noqa: E402


def main():
    eng = lrc.RaycastEngineGPU()
    mesh = lrc.synthetic.office()
    eng.set_mesh(mesh)
    ctx = eng.ctx
    poses = lrc.poses_from_waypoints(lrc.synthetic.office_waypoints(60))
    for name, intr in (("8-line", lrc.Indoor8LineLidarIntrinsics.create_standard_8line()),
                       ("32-line", lrc.Indoor8LineLidarIntrinsics.create_dense_32line())):
        for p in poses[:5]:
            eng.lidar_intersect_mesh(lrc.create_lidar(intr, p), mesh)
        t = {k: [] for k in ("create_lidar", "engine_call", "library_call", "copies", "device_only")}
        n = lrc.rays_per_frame(intr)
        st = (torch.empty((n, 3), dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory(),
              torch.zeros(2, dtype=torch.int64).pin_memory())
        for p in poses[5:]:
            t0 = time.perf_counter()
            lidar = lrc.create_lidar(intr, p)
            t1 = time.perf_counter()
            pts, inc = eng.lidar_intersect_mesh(lidar, mesh)
            t2 = time.perf_counter()
            m = len(pts)
            a, b = st[0][:m].numpy().copy(), st[1][:m].numpy().copy()
            t3 = time.perf_counter()
            t["create_lidar"].append(t1 - t0); t["engine_call"].append(t2 - t1); t["copies"].append(t3 - t2)
            # the library call alone (descriptor reuse, no copies out)
            d = core.single_axis_desc(intr)
            out = core.nat.Out(C.c_void_p(st[0].data_ptr()), C.c_void_p(st[1].data_ptr()), None, None, None, C.c_void_p(st[2].data_ptr()), int(st[0].shape[0]))
            total = C.c_int64(0)
            ph = np.ascontiguousarray(p, dtype=np.float64).reshape(16)
            t4 = time.perf_counter()
            core.nat.check(ctx._h, ctx._lib.lrc_scan_single_axis_host(ctx._h, C.c_void_p(ph.ctypes.data), 1, C.byref(d), None, C.byref(out), 1, C.byref(total)))
            t5 = time.perf_counter()
            t["library_call"].append(t5 - t4)
            # kernels only: device-resident scan of the same frame, CUDA events
            pd = torch.from_numpy(ph.reshape(1, 16)).cuda()
            bufs, _ = ctx._alloc_out(lrc.rays_per_frame(intr), 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.scan_enqueue(pd, intr, None, bufs); e1.record(); torch.cuda.synchronize()
            t["device_only"].append(e0.elapsed_time(e1) * 1e-3)
        print(name, {k: round(float(np.median(v)) * 1e6, 1) for k, v in t.items()}, "us (medians)", flush=True)


if __name__ == "__main__":
    main()
