"""
CPU ORACLE (Python side) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under the product package does.

It binds ``oracle/liblrc_oracle.so`` (plain C, see ``lrc_oracle.c``) with ctypes and restates,
in numpy and in the reference's own order of operations, the frame-level wrapper of

    /root/reference/raycast_engine/raycast_engine_cpu.py:24-111   (RaycastEngineCPU)

with the Open3D/Embree ``RaycastingScene`` (absent here, see lrc_oracle.c header) replaced by
the C intersector.  Like the reference, ``OracleEngineCPU`` builds a NEW scene on every call
(raycast_engine_cpu.py:46-47); ``OracleScene`` exposes the prebuilt-scene variant used for the
"cast only" CPU figure.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblrc_oracle.so")
MISS_ID = 0xFFFFFFFF


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc, see oracle/Makefile)."""
    src = os.path.join(_HERE, "lrc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class DualParams(C.Structure):
    _fields_ = [
        ("num_lines", C.c_int32),
        ("points_per_line", C.c_int32),
        ("theta_min", C.c_double),
        ("theta_max", C.c_double),
        ("swing_amplitude", C.c_double),
        ("swing_frequency", C.c_double),
        ("angle_noise_std", C.c_double),
        ("dropout_probability", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, i32, u64, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_double
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [vp, i64, vp, i64]
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_cast_rays.argtypes = [vp, vp, i64, vp, vp]
        L.orc_scene_stats.argtypes = [vp, vp, vp, vp, vp]
        L.orc_cast_rays_brute.argtypes = [vp, vp, i64, vp, i64, vp, vp]
        L.orc_cast_rays_brute_f64.argtypes = [vp, vp, i64, vp, i64, vp, vp]
        L.orc_gen_rays_single_axis.argtypes = [vp, vp, i32, i32, vp]
        L.orc_gen_rays_uniform.argtypes = [vp, dbl, dbl, i32, i32, vp]
        L.orc_gen_rays_dual_axis.argtypes = [vp, C.POINTER(DualParams), u64, u64, vp, vp, vp]
        L.orc_philox.argtypes = [u64, u64, C.c_uint32, C.c_uint32, vp]
        L.orc_epilogue.restype = i64
        L.orc_epilogue.argtypes = [vp, vp, vp, vp, i64, vp, dbl, vp, vp, vp, vp, vp, vp]
        L.orc_num_threads.restype = i32
        L.orc_set_num_threads.argtypes = [i32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


# ------------------------------------------------------------------------------------------------
# mesh access: Open3D legacy TriangleMesh duck type (``.vertices`` V x 3 float64, ``.triangles``
# T x 3 int32) or a (V, F) pair.  Vertices are rounded to float32 exactly as
# o3d.t.geometry.TriangleMesh.from_legacy does (raycast_engine_cpu.py:47).
# ------------------------------------------------------------------------------------------------
def mesh_arrays(mesh):
    if isinstance(mesh, (tuple, list)):
        v, f = mesh[0], mesh[1]
    else:
        v, f = mesh.vertices, mesh.triangles
    v = np.ascontiguousarray(np.asarray(v), dtype=np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(np.asarray(f), dtype=np.int32).reshape(-1, 3)
    return v, f


class OracleScene:
    """A prebuilt CPU scene (binned-SAH BVH2 over float32 triangles)."""

    def __init__(self, mesh):
        self.verts, self.tris = mesh_arrays(mesh)
        self._h = lib().orc_scene_create(_p(self.verts), len(self.verts), _p(self.tris), len(self.tris))
        if not self._h:
            raise MemoryError("orc_scene_create failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def cast_rays(self, rays: np.ndarray):
        """-> (t_hit float32 [N] with +inf on miss, prim_id uint32 [N] with 0xFFFFFFFF on miss)."""
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        n = len(rays)
        t = np.empty(n, np.float32)
        pid = np.empty(n, np.uint32)
        lib().orc_cast_rays(self._h, _p(rays), n, _p(t), _p(pid))
        return t, pid

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        d = C.c_int32()
        lib().orc_scene_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return {"box_tests": a.value, "tri_tests": b.value, "rays": c.value, "n_nodes": d.value}


def cast_rays_brute(mesh, rays):
    v, f = mesh_arrays(mesh)
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    t = np.empty(len(rays), np.float32)
    pid = np.empty(len(rays), np.uint32)
    lib().orc_cast_rays_brute(_p(v), _p(f), len(f), _p(rays), len(rays), _p(t), _p(pid))
    return t, pid


def cast_rays_brute_f64(mesh, rays):
    v, f = mesh_arrays(mesh)
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    t = np.empty(len(rays), np.float64)
    pid = np.empty(len(rays), np.uint32)
    lib().orc_cast_rays_brute_f64(_p(v), _p(f), len(f), _p(rays), len(rays), _p(t), _p(pid))
    return t, pid


# ------------------------------------------------------------------------------------------------
# ray tables
# ------------------------------------------------------------------------------------------------
def _pose(pose):
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    assert pose.shape == (4, 4)
    return pose


def gen_rays_single_axis(pose, vertical_degrees, W: int) -> np.ndarray:
    """lidar/indoor_lidar.py:94-131."""
    pose = _pose(pose)
    W = max(1, int(W))
    vd = np.ascontiguousarray(vertical_degrees if vertical_degrees else [0.0], dtype=np.float64)
    rays = np.empty((len(vd) * W, 6), np.float32)
    lib().orc_gen_rays_single_axis(_p(pose), _p(vd), len(vd), W, _p(rays))
    return rays


def gen_rays_uniform(pose, fov_up, fov_down, H: int, W: int) -> np.ndarray:
    """lidar/indoor_lidar.py:56-91."""
    pose = _pose(pose)
    H, W = max(1, int(H)), max(1, int(W))
    rays = np.empty((H * W, 6), np.float32)
    lib().orc_gen_rays_uniform(_p(pose), float(fov_up), float(fov_down), H, W, _p(rays))
    return rays


def dual_params(intr, angle_noise_std=None, dropout_probability=None) -> DualParams:
    """Field mapping of DualAxisLidarIntrinsics (lidar/lidar_intrinsics.py:28-66), duck-typed."""
    n_pts = int(intr.point_rate * intr.scan_duration)            # indoor_lidar.py:241
    return DualParams(
        num_lines=int(intr.num_vertical_lines),
        points_per_line=n_pts // int(intr.num_vertical_lines),   # indoor_lidar.py:244
        theta_min=float(intr.theta_range[0]),
        theta_max=float(intr.theta_range[1]),
        swing_amplitude=float(intr.swing_amplitude),
        swing_frequency=float(intr.swing_frequency),
        angle_noise_std=float(intr.angle_noise_std if angle_noise_std is None else angle_noise_std),
        dropout_probability=float(intr.dropout_probability if dropout_probability is None else dropout_probability),
    )


def gen_rays_dual_axis(pose, params: DualParams, seed: int = 0, pose_idx: int = 0, compact: bool = True):
    """lidar/indoor_lidar.py:224-296.  Returns rays (dropped rays removed when ``compact``) and the
    dense keep mask."""
    pose = _pose(pose)
    n = params.num_lines * params.points_per_line
    rays = np.empty((n, 6), np.float32)
    keep = np.empty(n, np.uint8)
    kept = C.c_int64()
    lib().orc_gen_rays_dual_axis(_p(pose), C.byref(params), int(seed), int(pose_idx), _p(rays), _p(keep), C.byref(kept))
    mask = keep.astype(bool)
    if compact:
        return rays[mask], mask
    return rays, mask


def philox(seed: int, pose_idx: int, ray_idx: int, stream: int = 0) -> np.ndarray:
    out = np.empty(4, np.uint32)
    lib().orc_philox(int(seed), int(pose_idx), int(ray_idx), int(stream), _p(out))
    return out


# ------------------------------------------------------------------------------------------------
# frame epilogue in C (used for the multi-threaded "cast only" CPU figure and as a second
# implementation the numpy restatement below is checked against)
# ------------------------------------------------------------------------------------------------
@dataclass
class Frame:
    points: np.ndarray      # (M,3) float32
    incident: np.ndarray    # (M,)  float64, degrees
    prim_id: np.ndarray     # (M,)  uint32
    label: np.ndarray       # (M,)  uint32   sem | ins << 16
    ray_idx: np.ndarray     # (M,)  uint32   index of the ray inside the (dense) frame


def epilogue_c(rays, t_hit, prim_id, center=None, max_range=-1.0, tri_label=None, keep=None) -> Frame:
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    n = len(rays)
    xyz = np.empty((n, 3), np.float32)
    inc = np.empty(n, np.float64)
    op = np.empty(n, np.uint32)
    ol = np.empty(n, np.uint32)
    orr = np.empty(n, np.uint32)
    c = np.zeros(3, np.float64) if center is None else np.ascontiguousarray(center, dtype=np.float64)
    lab = None if tri_label is None else np.ascontiguousarray(tri_label, dtype=np.uint32)
    kp = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    m = lib().orc_epilogue(_p(rays), _p(np.ascontiguousarray(t_hit, np.float32)),
                           _p(np.ascontiguousarray(prim_id, np.uint32)), _p(kp), n, _p(c), float(max_range),
                           _p(lab), _p(xyz), _p(inc), _p(op), _p(ol), _p(orr))
    return Frame(xyz[:m].copy(), inc[:m].copy(), op[:m].copy(), ol[:m].copy(), orr[:m].copy())


# ------------------------------------------------------------------------------------------------
# numpy restatement of RaycastEngineCPU, statement by statement
# ------------------------------------------------------------------------------------------------
class OracleEngineCPU:
    """Behavioural twin of /root/reference/raycast_engine/raycast_engine_cpu.py:16-111."""

    def __init__(self):
        self.last_prim_id = None   # extra: ids of the returned points (the reference discards them)

    def rays_intersect_mesh(self, rays: np.ndarray, mesh):
        if not isinstance(rays, np.ndarray):                       # :40-41
            raise TypeError("rays must be a numpy array.")
        if rays.ndim != 2 or rays.shape[1] != 6:                   # :42-43
            raise ValueError("rays must be a (N, 6) array.")
        scene = OracleScene(mesh)                                  # :46-47  new scene per call
        try:
            rays = rays.astype(np.float32)                         # :50
            depths, prim = scene.cast_rays(rays)                   # :51,:53
        finally:
            scene.close()
        masks = depths != np.inf                                   # :54
        rays_o = rays[:, :3]
        rays_d = rays[:, 3:]
        rays_d = rays_d / np.linalg.norm(rays_d, axis=1, keepdims=True)   # :57
        valid = np.isfinite(depths)                                # :60
        points = np.zeros_like(rays_o)
        points[valid] = rays_o[valid] + rays_d[valid] * depths[valid, None]   # :62
        self.last_prim_id = prim[masks]
        return points[masks]                                       # :71

    def lidar_intersect_mesh(self, lidar, mesh):
        rays = lidar.get_rays()                                    # :91
        points = self.rays_intersect_mesh(mesh=mesh, rays=rays)    # :92
        center = lidar.pose[:3, 3]                                 # :95
        dists = np.linalg.norm(points - center, axis=1)            # :96
        sel = dists < lidar.intrinsics.max_range                   # :97
        points = points[sel]
        self.last_prim_id = self.last_prim_id[sel]
        if len(points) > 0:                                        # :100-107
            directions = points - center
            directions = directions / np.linalg.norm(directions, axis=1, keepdims=True)
            incident = np.degrees(np.arccos(np.abs(directions[:, 2])))
        else:
            incident = np.empty(0)                                 # :109
        return points, incident
