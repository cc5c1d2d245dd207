"""Timing breakdown of the end-to-end path (run on the GPU box)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lrc_b200 as lrc

def t(fn, n=5):
    torch.cuda.synchronize(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

dev = torch.device("cuda", 0)
mesh = lrc.synthetic.office()
poses = lrc.poses_from_waypoints(lrc.synthetic.office_waypoints(100))
intr = lrc.Indoor8LineLidarIntrinsics.create_dense_32line()
eng = lrc.RaycastEngineGPU(); ctx = eng.ctx
v, f, l = lrc.mesh_arrays(mesh)
pv, pf, pl = torch.from_numpy(v).pin_memory(), torch.from_numpy(f).pin_memory(), torch.from_numpy(l.view(np.int32)).pin_memory()
big = torch.empty(307_200_000, dtype=torch.uint8, device=dev); hbig = torch.empty(307_200_000, dtype=torch.uint8).pin_memory()
print("D2H 307MB pinned       ms", t(lambda: hbig.copy_(big, non_blocking=True)))
print("H2D 307MB pinned       ms", t(lambda: big.copy_(hbig, non_blocking=True)))
print("mesh H2D (22MB)        ms", t(lambda: (pv.to(dev, non_blocking=True), pf.to(dev, non_blocking=True), pl.to(dev, non_blocking=True))))
vd, fd, ld = pv.to(dev), pf.to(dev), pl.to(dev)
print("set_mesh (resident)    ms", t(lambda: ctx.set_mesh_arrays(vd, fd, ld)))
pd = torch.from_numpy(poses.reshape(-1, 16)).to(dev)
bufs, _ = ctx._alloc_out(100 * 128000, 100)
print("scan (resident)        ms", t(lambda: ctx.scan_enqueue(pd, intr, None, bufs)))
host = ctx.alloc_host_buffers(100 * 128000, 100)
pp = torch.from_numpy(poses.reshape(-1, 16)).pin_memory()
for c in (2, 5, 10, 25, 100):
    print(f"scan_to_host chunk={c:3d}   ms", t(lambda: ctx.scan_to_host(pp, intr, None, host=host, chunk_poses=c)))
m = 12_799_947
def plain():
    ctx.scan_enqueue(pd, intr, None, bufs)
    n = int(bufs["off"][-1].item())
    host["points"][:n].copy_(bufs["xyz"][:n], non_blocking=True)
    host["incident"][:n].copy_(bufs["incident"][:n], non_blocking=True)
    host["label"][:n].copy_(bufs["label"][:n], non_blocking=True)
print("scan + 3 D2H unpipelined ms", t(plain))
print("3 D2H only             ms", t(lambda: (host["points"][:m].copy_(bufs["xyz"][:m], non_blocking=True), host["incident"][:m].copy_(bufs["incident"][:m], non_blocking=True), host["label"][:m].copy_(bufs["label"][:m], non_blocking=True))))
print("set_mesh_host          ms", t(lambda: ctx.set_mesh_host(pv, pf, pl)))
