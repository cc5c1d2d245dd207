"""
Host-side plumbing between the reference-shaped Python API and the C ABI (include/lrc.h).

PyTorch is used only for what the task allows it for: device memory (tensors own every buffer that
crosses the ABI), streams and (in ``distributed.py``) ``torch.distributed``.  All geometry, ray
generation, traversal and compaction happens inside ``csrc/liblrc.so``.
"""
from __future__ import annotations

import ctypes as C
import zlib
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat

MISS_ID = nat.MISS_ID


# --------------------------------------------------------------------------------------------------
# mesh access
# --------------------------------------------------------------------------------------------------
class TriangleMesh:
    """Minimal stand-in for ``o3d.geometry.TriangleMesh`` (Open3D is not a dependency of this engine):
    ``vertices`` (V,3) float64, ``triangles`` (T,3) int32 and, as an extension, ``triangle_labels`` (T,)
    uint32 = semantic | instance << 16.  Any object with ``.vertices`` / ``.triangles`` works as well."""

    def __init__(self, vertices, triangles, triangle_labels=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.triangles = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
        self.triangle_labels = None if triangle_labels is None else np.ascontiguousarray(triangle_labels, dtype=np.uint32)
        if self.triangle_labels is not None and len(self.triangle_labels) != len(self.triangles):
            raise ValueError("triangle_labels must have one entry per triangle")

    def __repr__(self):
        return f"TriangleMesh(V={len(self.vertices)}, T={len(self.triangles)}, labels={self.triangle_labels is not None})"


def pack_labels(semantic, instance) -> np.ndarray:
    """uint32 = semantic | instance << 16 (uint16/uint16 contract of reference s3dis_sim_scene.py:581-582)."""
    return (np.asarray(semantic, dtype=np.uint32) & 0xFFFF) | ((np.asarray(instance, dtype=np.uint32) & 0xFFFF) << 16)


def unpack_labels(label) -> Tuple[np.ndarray, np.ndarray]:
    label = np.asarray(label, dtype=np.uint32)
    return (label & 0xFFFF).astype(np.uint16), (label >> 16).astype(np.uint16)


def mesh_arrays(mesh):
    """-> (verts float32 (V,3), tris int32 (T,3), labels uint32 (T,) | None).  Vertices are rounded to
    float32 exactly as ``o3d.t.geometry.TriangleMesh.from_legacy`` does (reference raycast_engine_cpu.py:47)."""
    if isinstance(mesh, (tuple, list)):
        v, f = mesh[0], mesh[1]
        lab = mesh[2] if len(mesh) > 2 else None
    else:
        v, f = mesh.vertices, mesh.triangles
        lab = getattr(mesh, "triangle_labels", None)
    v = np.ascontiguousarray(np.asarray(v), dtype=np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(np.asarray(f), dtype=np.int32).reshape(-1, 3)
    if lab is not None:
        lab = np.ascontiguousarray(np.asarray(lab), dtype=np.uint32).reshape(-1)
        if len(lab) != len(f):
            raise ValueError("triangle_labels must have one entry per triangle")
    return v, f, lab


_CRC_POOL = None
_CRC_PIECE = 1 << 21          # bytes per checksum task


def _crc_pieces(arr) -> tuple:
    """CRC-32 of every 2 MiB piece of a buffer, computed on a small thread pool (zlib releases the GIL)."""
    global _CRC_POOL
    mv = memoryview(np.ascontiguousarray(arr)).cast("B")
    n = len(mv)
    if n <= _CRC_PIECE:
        return (zlib.crc32(mv),)
    if _CRC_POOL is None:
        import os
        from concurrent.futures import ThreadPoolExecutor
        try:
            workers = len(os.sched_getaffinity(0))
        except Exception:
            workers = os.cpu_count() or 1
        _CRC_POOL = ThreadPoolExecutor(max_workers=max(1, min(8, workers)))
    return tuple(_CRC_POOL.map(zlib.crc32, [mv[a:a + _CRC_PIECE] for a in range(0, n, _CRC_PIECE)]))


def mesh_fingerprint(mesh) -> tuple:
    """Content identity of a mesh for the BVH cache: sizes, dtypes and CRC-32s of the COMPLETE vertex, index and label
    buffers (2 MiB pieces on up to 8 threads: about 1-2 ms per million triangles, 6 ms on one core).  The reference rebuilds
    its scene on every call (raycast_engine_cpu.py:46-47), so reusing a BVH is only legal when nothing changed -- an
    in-place edit of one vertex or one label must be seen, which a sampled checksum cannot promise.  Callers that want to
    skip even this pass pin the mesh explicitly (``Context.pin_mesh`` / ``RaycastEngineGPU.set_mesh``)."""
    if isinstance(mesh, (tuple, list)):
        v, f = np.asarray(mesh[0]), np.asarray(mesh[1])
        lab = mesh[2] if len(mesh) > 2 else None
    else:
        v, f = np.asarray(mesh.vertices), np.asarray(mesh.triangles)
        lab = getattr(mesh, "triangle_labels", None)
    crc = _crc_pieces(v) + _crc_pieces(f)
    n_lab = -1
    if lab is not None:
        lab = np.asarray(lab)
        n_lab = int(lab.shape[0])
        crc = crc + _crc_pieces(lab)
    return (v.shape, str(v.dtype), f.shape, str(f.dtype), n_lab, crc)


# --------------------------------------------------------------------------------------------------
# sensor descriptors
# --------------------------------------------------------------------------------------------------
def single_axis_desc(intr) -> nat.SingleAxis:
    """Indoor8LineLidarIntrinsics -> lrc_single_axis (duck-typed; reference indoor_lidar.py:38-51)."""
    vd = getattr(intr, "vertical_degrees", None)
    d = nat.SingleAxis()
    d.W = max(1, int(intr.horizontal_res))
    if vd is None:
        d.H = max(1, int(intr.vertical_res))
        d.h_vertical_deg = None
        d._keep = None
    else:
        table = list(vd) if len(vd) > 0 else [0.0]          # reference indoor_lidar.py:104-106
        arr = (C.c_double * len(table))(*[float(x) for x in table])
        d.H = len(table)
        d.h_vertical_deg = C.cast(arr, C.POINTER(C.c_double))
        d._keep = arr                                         # keep the host table alive
    d.fov_up_deg = float(intr.fov_up)
    d.fov_down_deg = float(intr.fov_down)
    d.max_range = float(intr.max_range)
    return d


def dual_axis_desc(intr) -> nat.DualAxis:
    """DualAxisLidarIntrinsics -> lrc_dual_axis (reference indoor_lidar.py:241-252)."""
    n_pts = int(intr.point_rate * intr.scan_duration)
    lines = int(intr.num_vertical_lines)
    return nat.DualAxis(num_lines=lines, points_per_line=n_pts // lines,
                        theta_min=float(intr.theta_range[0]), theta_max=float(intr.theta_range[1]),
                        swing_amplitude=float(intr.swing_amplitude), swing_frequency=float(intr.swing_frequency),
                        max_range=float(intr.max_range))


def is_dual_axis(intr) -> bool:
    return hasattr(intr, "num_vertical_lines") and hasattr(intr, "swing_amplitude")


def rays_per_frame(intr) -> int:
    if is_dual_axis(intr):
        d = dual_axis_desc(intr)
        return d.num_lines * d.points_per_line
    d = single_axis_desc(intr)
    return d.H * d.W


@dataclass
class NoiseConfig:
    """Counter-based (Philox) noise; all zero = disabled = the bit-exact parity configuration."""
    angle_noise_std: float = 0.0
    dropout_probability: float = 0.0
    range_noise_std: float = 0.0
    seed: int = 0
    pose_index_base: int = 0

    def struct(self) -> nat.Noise:
        return nat.Noise(float(self.angle_noise_std), float(self.dropout_probability), float(self.range_noise_std),
                         int(self.seed) & 0xFFFFFFFFFFFFFFFF, int(self.pose_index_base))

    @classmethod
    def from_intrinsics(cls, intr, seed: int = 0, pose_index_base: int = 0) -> "NoiseConfig":
        """The noise the reference applies on the simulation path: angle noise + dropout for the dual-axis
        sensor (indoor_lidar.py:270-272,292-294); nothing for the single-axis sensor (add_noise is dead code)."""
        if is_dual_axis(intr):
            return cls(float(intr.angle_noise_std), float(intr.dropout_probability), 0.0, seed, pose_index_base)
        return cls(0.0, 0.0, 0.0, seed, pose_index_base)


# --------------------------------------------------------------------------------------------------
# results
# --------------------------------------------------------------------------------------------------
@dataclass
class ScanResult:
    """Compacted, ray-ordered hits of P frames (device tensors).  Frame p is [frame_offset[p], frame_offset[p+1])."""
    points: torch.Tensor          # (M,3) float32
    incident: torch.Tensor        # (M,)  float64 degrees
    prim_id: torch.Tensor         # (M,)  int32 holding uint32 bits
    label: torch.Tensor           # (M,)  int32 holding uint32 bits (sem | ins << 16)
    ray_idx: torch.Tensor         # (M,)  int32
    frame_offset: torch.Tensor    # (P+1,) int64, host copy in frame_offset_host
    frame_offset_host: np.ndarray

    @property
    def num_points(self) -> int:
        return int(self.frame_offset_host[-1])

    @property
    def num_frames(self) -> int:
        return len(self.frame_offset_host) - 1

    def frame(self, p: int):
        """(points (m,3) f32, incident (m,) f64) of frame p as numpy -- the reference's return value."""
        a, b = int(self.frame_offset_host[p]), int(self.frame_offset_host[p + 1])
        return self.points[a:b].cpu().numpy(), self.incident[a:b].cpu().numpy()

    def numpy(self) -> dict:
        return {
            "points": self.points.cpu().numpy(),
            "incident": self.incident.cpu().numpy(),
            "prim_id": self.prim_id.cpu().numpy().view(np.uint32),
            "label": self.label.cpu().numpy().view(np.uint32),
            "ray_idx": self.ray_idx.cpu().numpy().view(np.uint32),
            "frame_offset": self.frame_offset_host.copy(),
        }


# --------------------------------------------------------------------------------------------------
# page-locked result buffers that are handed out as numpy arrays
# --------------------------------------------------------------------------------------------------
class _PinnedLease:
    """One page-locked buffer on loan.  ``array`` wraps it as a numpy array whose base keeps this lease alive; when the last
    such array is dropped the buffer returns to its pool (CPython reference counting makes that immediate)."""
    __slots__ = ("pool", "tensor", "ptr")

    def __init__(self, pool, tensor):
        self.pool, self.tensor, self.ptr = pool, tensor, tensor.data_ptr()

    def array(self, dtype, count: int) -> np.ndarray:
        buf = (C.c_char * (np.dtype(dtype).itemsize * count)).from_address(self.ptr)
        buf._lease = self                       # ndarray -> memoryview -> ctypes array -> lease -> pinned tensor
        return np.frombuffer(buf, dtype=dtype, count=count)

    def __del__(self):
        pool = self.pool
        if pool is not None and pool.free is not None:
            pool.free.append(self.tensor)


class _PinnedPool:
    """At most ``max_items`` page-locked buffers of ``item_bytes`` each, allocated on demand (cudaHostAlloc is slow: a miss
    costs about a millisecond, a hit nothing)."""

    def __init__(self, item_bytes: int, max_items: int = 8):
        self.item_bytes, self.max_items, self.count = int(item_bytes), int(max_items), 0
        self.free: Optional[list] = []

    def acquire(self) -> Optional[_PinnedLease]:
        if self.free:
            return _PinnedLease(self, self.free.pop())
        if self.count >= self.max_items:
            return None
        self.count += 1
        return _PinnedLease(self, torch.empty(self.item_bytes, dtype=torch.uint8).pin_memory())


# --------------------------------------------------------------------------------------------------
# context
# --------------------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Context:
    """One ``lrc_ctx`` (one GPU).  Not thread-safe; all work goes to torch's current stream."""

    def __init__(self, device: Optional[int] = None):
        self._lib = nat.load()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device visible: the LiDAR ray-casting engine has no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        nat.check(None, self._lib.lrc_create(self.device_index, C.byref(h)))
        self._h = h
        self._mesh_key = None
        self._pinned = None               # the mesh object whose BVH is resident and trusted without hashing (pin_mesh)
        self.num_tris = 0
        self.has_labels = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.lrc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ----
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, arr, dtype) -> torch.Tensor:
        if isinstance(arr, torch.Tensor):
            return arr.to(device=self.device, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(arr)).to(device=self.device, dtype=dtype, non_blocking=False).contiguous()

    def launch_count(self) -> int:
        return int(self._lib.lrc_launch_count(self._h))

    def set_option(self, key: str, value: int) -> None:
        nat.check(self._h, self._lib.lrc_set_option(self._h, key.encode(), int(value)))

    def stat(self, key: str) -> int:
        """``lrc_get_stat``: scratch_bytes, bvh_bytes, node_format, build_quality, ploc_iterations and the generation
        counters of the context's single mesh / NN index / collision index slots."""
        v = C.c_int64(0)
        nat.check(self._h, self._lib.lrc_get_stat(self._h, key.encode(), C.byref(v)))
        return int(v.value)

    def default_l2_persist(self) -> int:
        """The library's default for option ``l2_persist`` (percent of the maximum persisting-L2 set-aside)."""
        return int(self._lib.lrc_default_l2_persist())

    def kernel_times(self) -> dict:
        """Device time (ms) of k_trace and of the compaction kernels of the last scan (option ``kernel_timing``)."""
        tr, cp, n = C.c_double(0), C.c_double(0), C.c_int32(0)
        nat.check(self._h, self._lib.lrc_kernel_times(self._h, C.byref(tr), C.byref(cp), C.byref(n)))
        return {"trace_ms": tr.value, "compact_ms": cp.value, "trace_launches": n.value}

    # ---- scene ----
    def set_mesh_arrays(self, verts, tris, labels=None) -> None:
        """Upload float32 vertices / int32 indices / uint32 labels and build the LBVH on the GPU."""
        self._mesh_key = None              # whatever happens below, the resident BVH is no longer the cached mesh's
        self._pinned = None
        with torch.cuda.device(self.device):
            v = self._dev(verts, torch.float32).reshape(-1, 3)
            f = self._dev(tris, torch.int32).reshape(-1, 3)
            lab = None
            if labels is not None:
                lab_np = labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(labels, dtype=np.uint32).view(np.int32))
                lab = lab_np.to(device=self.device).contiguous()
                if lab.numel() != f.shape[0]:
                    raise ValueError("labels must have one entry per triangle")
            nat.check(self._h, self._lib.lrc_set_mesh(self._h, _ptr(v), v.shape[0], _ptr(f), f.shape[0], _ptr(lab), self._stream()))
            self.num_tris = int(f.shape[0])
            self.has_labels = lab is not None

    def set_mesh(self, mesh, cache: bool = True) -> bool:
        """Build (or reuse) the BVH of ``mesh``.  Returns True when a build happened.

        ``cache=True`` reuses the resident BVH when the mesh is the pinned object (no hashing, see ``pin_mesh``) or when
        its full-content fingerprint equals the resident one; ``cache=False`` always rebuilds, like the reference."""
        if cache and self._pinned is not None and mesh is self._pinned:
            return False
        key = mesh_fingerprint(mesh) if cache else None
        if cache and self._mesh_key is not None and key == self._mesh_key:
            return False
        v, f, lab = mesh_arrays(mesh)
        self.set_mesh_arrays(v, f, lab)
        self._mesh_key = key
        return True

    def pin_mesh(self, mesh) -> bool:
        """Build the BVH of ``mesh`` (unless its content is already resident) and trust this OBJECT from now on: later
        ``set_mesh(mesh)`` calls with the same object return at once, without reading its arrays.  The caller promises
        not to edit the arrays in place; ``unpin_mesh()`` (or any other mesh) ends the promise."""
        self._pinned = None
        built = self.set_mesh(mesh, cache=True)
        self._pinned = mesh
        return built

    def unpin_mesh(self) -> None:
        self._pinned = None

    def invalidate_mesh(self) -> None:
        """Forget the pinned object and the cached fingerprint: the next ``set_mesh`` rebuilds (needed after changing a
        build option such as ``leaf_size`` / ``node_format`` / ``build_quality``)."""
        self._pinned = None
        self._mesh_key = None

    def bvh_info(self) -> dict:
        info = nat.BvhInfo()
        nat.check(self._h, self._lib.lrc_bvh_get_info(self._h, C.byref(info)))
        return {
            "num_tris": info.num_tris, "num_nodes": info.num_nodes, "max_depth": info.max_depth,
            "scene_min": list(info.scene_min), "scene_max": list(info.scene_max), "box_pad": info.box_pad,
            "sah_cost": info.sah_cost, "bytes_nodes": info.bytes_nodes, "bytes_tris": info.bytes_tris,
        }

    # ---- counters ----
    def set_counting(self, enabled: bool) -> None:
        nat.check(self._h, self._lib.lrc_set_counting(self._h, 1 if enabled else 0))

    def counters(self, reset: bool = True) -> dict:
        c = nat.Counters()
        nat.check(self._h, self._lib.lrc_counters(self._h, C.byref(c), 1 if reset else 0, self._stream()))
        return {"rays": c.rays, "nodes_visited": c.nodes_visited, "tris_tested": c.tris_tested, "hits": c.hits}

    # ---- rays_intersect_mesh family ----
    def cast_rays(self, rays, bruteforce: bool = False):
        """rays (N,6) float32 (tensor or ndarray) -> (t_hit float32 [N], prim_id int32 [N] holding uint32 bits)."""
        with torch.cuda.device(self.device):
            r = self._dev(rays, torch.float32).reshape(-1, 6)
            n = r.shape[0]
            t = torch.empty(n, dtype=torch.float32, device=self.device)
            pid = torch.empty(n, dtype=torch.int32, device=self.device)
            fn = self._lib.lrc_cast_rays_bruteforce if bruteforce else self._lib.lrc_cast_rays
            nat.check(self._h, fn(self._h, _ptr(r), n, _ptr(t), _ptr(pid), self._stream()))
            return t, pid

    def _alloc_out(self, capacity: int, frames: int, full: bool = True):
        dev = self.device
        cap = max(1, capacity)
        bufs = {
            "xyz": torch.empty((cap, 3), dtype=torch.float32, device=dev),
            "incident": torch.empty(cap, dtype=torch.float64, device=dev) if full else None,
            "prim": torch.empty(cap, dtype=torch.int32, device=dev) if full else None,
            "label": torch.empty(cap, dtype=torch.int32, device=dev) if full else None,
            "ray": torch.empty(cap, dtype=torch.int32, device=dev) if full else None,
            "off": torch.zeros(frames + 1, dtype=torch.int64, device=dev),
        }
        out = nat.Out(_ptr(bufs["xyz"]), _ptr(bufs["incident"]), _ptr(bufs["prim"]), _ptr(bufs["label"]),
                      _ptr(bufs["ray"]), _ptr(bufs["off"]), cap)
        return bufs, out

    def _finish(self, bufs) -> ScanResult:
        off_host = bufs["off"].cpu().numpy()          # synchronises the stream
        m = int(off_host[-1])
        e32 = torch.empty(0, dtype=torch.int32, device=self.device)
        return ScanResult(
            points=bufs["xyz"][:m],
            incident=bufs["incident"][:m] if bufs["incident"] is not None else torch.empty(0, dtype=torch.float64, device=self.device),
            prim_id=bufs["prim"][:m] if bufs["prim"] is not None else e32,
            label=bufs["label"][:m] if bufs["label"] is not None else e32,
            ray_idx=bufs["ray"][:m] if bufs["ray"] is not None else e32,
            frame_offset=bufs["off"], frame_offset_host=off_host)

    def rays_intersect(self, rays) -> ScanResult:
        """== RaycastEngineCPU.rays_intersect_mesh core: hit points in ray order (no range filter)."""
        with torch.cuda.device(self.device):
            r = self._dev(rays, torch.float32).reshape(-1, 6)
            bufs, out = self._alloc_out(r.shape[0], 1)
            nat.check(self._h, self._lib.lrc_rays_intersect(self._h, _ptr(r), r.shape[0], C.byref(out), self._stream()))
            return self._finish(bufs)

    def scan_rays(self, rays, center, max_range: float) -> ScanResult:
        """== RaycastEngineCPU.lidar_intersect_mesh for explicit rays of one frame."""
        with torch.cuda.device(self.device):
            r = self._dev(rays, torch.float32).reshape(-1, 6)
            c = (C.c_double * 3)(*[float(x) for x in center])
            bufs, out = self._alloc_out(r.shape[0], 1)
            nat.check(self._h, self._lib.lrc_scan_rays(self._h, _ptr(r), r.shape[0], c, float(max_range), C.byref(out), self._stream()))
            return self._finish(bufs)

    # ---- batched trajectory scan ----
    def scan_enqueue(self, poses_dev: torch.Tensor, intr, noise: Optional[NoiseConfig], bufs=None):
        """Enqueue a P-pose scan on the current stream; returns (bufs, out struct) without synchronising."""
        P = int(poses_dev.shape[0])
        n_frame = rays_per_frame(intr)
        if bufs is None:
            bufs, out = self._alloc_out(P * n_frame, P)
        else:
            out = nat.Out(_ptr(bufs["xyz"]), _ptr(bufs["incident"]), _ptr(bufs["prim"]), _ptr(bufs["label"]),
                          _ptr(bufs["ray"]), _ptr(bufs["off"]), bufs["xyz"].shape[0])
        nz = noise.struct() if noise is not None else None
        nzp = C.byref(nz) if nz is not None else None
        if is_dual_axis(intr):
            d = dual_axis_desc(intr)
            nat.check(self._h, self._lib.lrc_scan_dual_axis(self._h, _ptr(poses_dev), P, C.byref(d), nzp, C.byref(out), self._stream()))
        else:
            d = single_axis_desc(intr)
            nat.check(self._h, self._lib.lrc_scan_single_axis(self._h, _ptr(poses_dev), P, C.byref(d), nzp, C.byref(out), self._stream()))
        return bufs, out

    def scan(self, poses, intr, noise: Optional[NoiseConfig] = None, bufs=None) -> ScanResult:
        """poses: (P,4,4) float64 (ndarray or tensor).  One fused launch sequence for the whole trajectory.
        ``bufs`` (from ``_alloc_out``) lets a caller reuse output buffers across calls; the returned tensors are views of
        them and are overwritten by the next call that uses the same buffers."""
        with torch.cuda.device(self.device):
            p = self._dev(poses, torch.float64).reshape(-1, 16)
            bufs, _ = self.scan_enqueue(p, intr, noise, bufs)
            return self._finish(bufs)

    def frame_to_numpy(self, res: "ScanResult"):
        """(points (m,3) float32, incident (m,) float64) of a one-frame result as fresh numpy arrays: both copies are
        enqueued into page-locked staging before the single synchronisation."""
        m = res.num_points
        if m == 0:
            return np.zeros((0, 3), np.float32), np.empty(0)
        st = getattr(self, "_frame_stage", None)
        if st is None or st[0].shape[0] < m:
            cap = max(m, 1 << 17)
            st = (torch.empty((cap, 3), dtype=torch.float32).pin_memory(), torch.empty(cap, dtype=torch.float64).pin_memory())
            self._frame_stage = st
        with torch.cuda.device(self.device):
            st[0][:m].copy_(res.points, non_blocking=True)
            st[1][:m].copy_(res.incident, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return st[0][:m].numpy().copy(), st[1][:m].numpy().copy()

    # ---- trajectory scan with HOST results: chunked, D2H overlapped with the next chunks' kernels ----
    def alloc_host_buffers(self, capacity: int, frames: int, labels: bool = True, extras: bool = False, incident: bool = True) -> dict:
        """Pinned host staging for ``scan_to_host`` (allocate once, reuse across calls).  ``incident=False`` leaves the
        incident angles on the device (8 of the 24 bytes per point that cross PCIe): they are a pure function of point and
        pose, the per-frame statistics that consume them run on the GPU (``frame_statistics``), and
        ``lrc_incident_angles`` reproduces them bit for bit from the points whenever they are wanted after all."""
        cap = max(1, int(capacity))
        h = {
            "points": torch.empty((cap, 3), dtype=torch.float32).pin_memory(),
            "incident": torch.empty(cap, dtype=torch.float64).pin_memory() if incident else None,
            "label": torch.empty(cap, dtype=torch.int32).pin_memory() if labels else None,
            "prim_id": torch.empty(cap, dtype=torch.int32).pin_memory() if extras else None,
            "ray_idx": torch.empty(cap, dtype=torch.int32).pin_memory() if extras else None,
            "frame_offset": torch.zeros(frames + 1, dtype=torch.int64).pin_memory(),
        }
        return h

    def scan_to_host(self, poses, intr, noise: Optional[NoiseConfig] = None, host: Optional[dict] = None,
                     chunk_poses: Optional[int] = None) -> dict:
        """``scan`` for callers that want numpy (``lrc_scan_*_host``): poses come from host memory, the trajectory is
        cut into pose chunks whose kernels are all enqueued up front, and each chunk's compacted records are copied
        into (pinned) host memory on a second stream as soon as that chunk is done -- the PCIe transfer, 24 B per
        point and the end-to-end bottleneck, overlaps the remaining chunks' traversal.  Synchronous."""
        poses_h = np.ascontiguousarray(poses.numpy() if isinstance(poses, torch.Tensor) else poses, dtype=np.float64).reshape(-1, 16)
        P = int(poses_h.shape[0])
        n_frame = rays_per_frame(intr)
        if host is None:
            host = self.alloc_host_buffers(P * n_frame, P)
        if host["points"].shape[0] < max(1, P * n_frame) or host["frame_offset"].shape[0] < P + 1:
            raise ValueError("host buffers are too small for this trajectory")

        def hp(name):
            t = host.get(name)
            return None if t is None else C.c_void_p(t.data_ptr())

        out = nat.Out(hp("points"), hp("incident"), hp("prim_id"), hp("label"), hp("ray_idx"), hp("frame_offset"),
                      int(host["points"].shape[0]))
        nz = noise.struct() if noise is not None else None
        nzp = C.byref(nz) if nz is not None else None
        total = C.c_int64(0)
        chunk = 0 if chunk_poses is None else int(chunk_poses)
        pp = C.c_void_p(poses_h.ctypes.data)
        with torch.cuda.device(self.device):
            if is_dual_axis(intr):
                d = dual_axis_desc(intr)
                nat.check(self._h, self._lib.lrc_scan_dual_axis_host(self._h, pp, P, C.byref(d), nzp, C.byref(out), chunk, C.byref(total)))
            else:
                d = single_axis_desc(intr)
                nat.check(self._h, self._lib.lrc_scan_single_axis_host(self._h, pp, P, C.byref(d), nzp, C.byref(out), chunk, C.byref(total)))
        m = int(total.value)
        res = {"points": host["points"][:m].numpy(), "frame_offset": host["frame_offset"][:P + 1].numpy(), "num_points": m}
        if host.get("incident") is not None:
            res["incident"] = host["incident"][:m].numpy()
        for k in ("label", "prim_id", "ray_idx"):
            if host.get(k) is not None:
                res[k] = host[k][:m].numpy().view(np.uint32)
        return res

    def scan_frame_to_host(self, pose, intr, noise: Optional[NoiseConfig] = None):
        """ONE frame, host in / host out, one library call and one synchronisation: the reference's per-waypoint call
        pattern (s3dis_simulator.py:254-263).  The pose goes up from host memory; points and incident angles are copied
        by the device straight into page-locked buffers that BECOME the returned numpy arrays (``_PinnedPool``: a buffer
        goes back to the pool when the caller drops the array, so a loop that consumes frames one by one never copies on
        the host; a caller that keeps every frame exhausts the pool and gets ordinary arrays filled from staging).
        ``lrc_scan_*_host`` copies a small frame at capacity right behind its kernels instead of waiting for the point
        count first.  -> (points (m,3) float32, incident (m,) float64)"""
        dual = is_dual_axis(intr)
        d = dual_axis_desc(intr) if dual else single_axis_desc(intr)
        n = max(1, d.num_lines * d.points_per_line if dual else d.H * d.W)
        pools = getattr(self, "_frame_pools", None)
        if pools is None or pools[0].item_bytes < 12 * n:
            pools = (_PinnedPool(12 * n), _PinnedPool(8 * n), torch.zeros(2, dtype=torch.int64).pin_memory())
            self._frame_pools = pools
        la, lb = pools[0].acquire(), pools[1].acquire()
        zero_copy = la is not None and lb is not None
        if not zero_copy:
            la = lb = None
            st = getattr(self, "_frame_host", None)
            if st is None or st[0].shape[0] < n:
                st = (torch.empty((n, 3), dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory())
                self._frame_host = st
            p_xyz, p_inc = st[0].data_ptr(), st[1].data_ptr()
        else:
            p_xyz, p_inc = la.ptr, lb.ptr
        pose_h = np.ascontiguousarray(pose, dtype=np.float64).reshape(16)
        out = nat.Out(C.c_void_p(p_xyz), C.c_void_p(p_inc), None, None, None, C.c_void_p(pools[2].data_ptr()), n)
        nz = noise.struct() if noise is not None else None
        total = C.c_int64(0)
        fn = self._lib.lrc_scan_dual_axis_host if dual else self._lib.lrc_scan_single_axis_host
        with torch.cuda.device(self.device):
            nat.check(self._h, fn(self._h, C.c_void_p(pose_h.ctypes.data), 1, C.byref(d), C.byref(nz) if nz is not None else None,
                                  C.byref(out), 1, C.byref(total)))
        m = int(total.value)
        if m == 0:
            return np.zeros((0, 3), np.float32), np.empty(0)
        if zero_copy:
            return la.array(np.float32, 3 * m).reshape(m, 3), lb.array(np.float64, m)
        return st[0][:m].numpy().copy(), st[1][:m].numpy().copy()

    def set_mesh_host(self, verts: np.ndarray, tris: np.ndarray, labels: Optional[np.ndarray] = None) -> None:
        """``lrc_set_mesh_host``: upload float32 vertices / int32 indices / uint32 labels from host memory (pinned
        tensors or numpy arrays) and build the LBVH.  Synchronous."""
        def as_np(a, dt):
            if isinstance(a, torch.Tensor):
                a = a.numpy()
            return np.ascontiguousarray(a, dtype=dt)
        self._mesh_key = None
        self._pinned = None
        v = as_np(verts, np.float32).reshape(-1, 3)
        f = as_np(tris, np.int32).reshape(-1, 3)
        lab = None if labels is None else as_np(labels, np.int32 if (isinstance(labels, torch.Tensor) or labels.dtype == np.int32) else np.uint32).reshape(-1)
        with torch.cuda.device(self.device):
            nat.check(self._h, self._lib.lrc_set_mesh_host(self._h, C.c_void_p(v.ctypes.data), v.shape[0], C.c_void_p(f.ctypes.data), f.shape[0],
                                                           None if lab is None else C.c_void_p(lab.ctypes.data)))
        self.num_tris = int(f.shape[0])
        self.has_labels = lab is not None

    # ---- get_rays ----
    def gen_rays(self, poses, intr, noise: Optional[NoiseConfig] = None):
        """-> (rays (P*N,6) float32 tensor, keep (P*N,) uint8 tensor | None)."""
        with torch.cuda.device(self.device):
            p = self._dev(poses, torch.float64).reshape(-1, 16)
            P = p.shape[0]
            n = P * rays_per_frame(intr)
            rays = torch.empty((max(n, 1), 6), dtype=torch.float32, device=self.device)[:n]
            if is_dual_axis(intr):
                d = dual_axis_desc(intr)
                keep = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)[:n]
                nz = noise.struct() if noise is not None else None
                nat.check(self._h, self._lib.lrc_gen_rays_dual_axis(self._h, _ptr(p), P, C.byref(d), C.byref(nz) if nz is not None else None,
                                                                     _ptr(rays), _ptr(keep), self._stream()))
                return rays, keep
            d = single_axis_desc(intr)
            nat.check(self._h, self._lib.lrc_gen_rays_single_axis(self._h, _ptr(p), P, C.byref(d), _ptr(rays), self._stream()))
            return rays, None


_contexts = {}


def get_context(device: Optional[int] = None) -> Context:
    """Process-wide context per GPU (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible: the LiDAR ray-casting engine has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else int(device)
    if idx not in _contexts:
        _contexts[idx] = Context(idx)
    return _contexts[idx]
