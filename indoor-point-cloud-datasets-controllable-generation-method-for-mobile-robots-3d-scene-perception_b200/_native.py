"""
ctypes binding of ``csrc/liblrc.so`` -- the C ABI declared in ``include/lrc.h``.

There is no fallback of any kind: if the shared object is missing, cannot be loaded, or no sm_100
device is present, the first call raises.  (Build with ``python __graft_entry__.py`` /
``<package>.build.build_native()``.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LRC_LIBRARY_PATH points at an alternative build of the same library (kernel experiments); there is still no fallback
LIB_PATH = os.environ.get("LRC_LIBRARY_PATH") or os.path.join(_HERE, "csrc", "liblrc.so")

MISS_ID = 0xFFFFFFFF


class LrcError(RuntimeError):
    """A liblrc call returned a negative status."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"liblrc error {code}: {msg}")
        self.code = code


class SingleAxis(C.Structure):          # lrc_single_axis
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("h_vertical_deg", C.POINTER(C.c_double)),
                ("fov_up_deg", C.c_double), ("fov_down_deg", C.c_double), ("max_range", C.c_double)]


class DualAxis(C.Structure):            # lrc_dual_axis
    _fields_ = [("num_lines", C.c_int32), ("points_per_line", C.c_int32), ("theta_min", C.c_double),
                ("theta_max", C.c_double), ("swing_amplitude", C.c_double), ("swing_frequency", C.c_double),
                ("max_range", C.c_double)]


class Noise(C.Structure):               # lrc_noise
    _fields_ = [("angle_noise_std", C.c_double), ("dropout_probability", C.c_double),
                ("range_noise_std", C.c_double), ("seed", C.c_uint64), ("pose_index_base", C.c_uint64)]


class Out(C.Structure):                 # lrc_out
    _fields_ = [("xyz", C.c_void_p), ("incident_deg", C.c_void_p), ("prim_id", C.c_void_p), ("label", C.c_void_p),
                ("ray_idx", C.c_void_p), ("frame_offset", C.c_void_p), ("capacity", C.c_int64)]


class IpcHandle(C.Structure):          # lrc_ipc_handle
    _fields_ = [("bytes", C.c_ubyte * 64)]


class Gather(C.Structure):             # lrc_gather
    _fields_ = [("n_targets", C.c_int32), ("frame_capacity", C.c_int32), ("xyz", C.c_void_p * 16), ("label", C.c_void_p * 16),
                ("frame_offset", C.c_void_p * 16), ("point_base", C.c_int64), ("frame_base", C.c_int64), ("capacity", C.c_int64)]


class Counters(C.Structure):            # lrc_counters_t
    _fields_ = [("rays", C.c_uint64), ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("hits", C.c_uint64)]


class FrameStats(C.Structure):          # lrc_frame_stats
    _fields_ = [("num_points", C.c_int64), ("incident_mean", C.c_double), ("incident_std", C.c_double),
                ("range_mean", C.c_double), ("range_std", C.c_double)]


class GatherWire(C.Structure):         # lrc_gather_wire
    _fields_ = [("enabled", C.c_int32), ("self_index", C.c_int32), ("t", C.c_void_p * 16), ("ray_idx", C.c_void_p * 16),
                ("ready", C.c_void_p * 16), ("all_poses", C.c_void_p), ("rank_pose0", C.c_int64 * 16), ("rank_frames", C.c_int64 * 16),
                ("rank_point_base", C.c_int64 * 16), ("rank_frame_base", C.c_int64 * 16)]


class BvhInfo(C.Structure):             # lrc_bvh_info
    _fields_ = [("num_tris", C.c_int64), ("num_nodes", C.c_int64), ("max_depth", C.c_int32), ("reserved", C.c_int32),
                ("scene_min", C.c_float * 3), ("scene_max", C.c_float * 3), ("box_pad", C.c_float),
                ("sah_cost", C.c_float), ("bytes_nodes", C.c_int64), ("bytes_tris", C.c_int64)]


# every symbol include/lrc.h declares: (name, restype, argtypes)
_vp, _i64, _i32, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
SYMBOLS = {
    "lrc_abi_version": (_i32, []),
    "lrc_create": (_i32, [_i32, C.POINTER(_vp)]),
    "lrc_destroy": (None, [_vp]),
    "lrc_last_error": (C.c_char_p, [_vp]),
    "lrc_set_mesh": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "lrc_bvh_get_info": (_i32, [_vp, C.POINTER(BvhInfo)]),
    "lrc_cast_rays": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "lrc_cast_rays_bruteforce": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "lrc_rays_intersect": (_i32, [_vp, _vp, _i64, C.POINTER(Out), _vp]),
    "lrc_scan_rays": (_i32, [_vp, _vp, _i64, C.POINTER(_dbl), _dbl, C.POINTER(Out), _vp]),
    "lrc_scan_single_axis": (_i32, [_vp, _vp, _i64, C.POINTER(SingleAxis), C.POINTER(Noise), C.POINTER(Out), _vp]),
    "lrc_scan_dual_axis": (_i32, [_vp, _vp, _i64, C.POINTER(DualAxis), C.POINTER(Noise), C.POINTER(Out), _vp]),
    "lrc_scan_single_axis_host": (_i32, [_vp, _vp, _i64, C.POINTER(SingleAxis), C.POINTER(Noise), C.POINTER(Out), _i64, C.POINTER(_i64)]),
    "lrc_scan_dual_axis_host": (_i32, [_vp, _vp, _i64, C.POINTER(DualAxis), C.POINTER(Noise), C.POINTER(Out), _i64, C.POINTER(_i64)]),
    "lrc_set_mesh_host": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "lrc_peer_buffer_create": (_i32, [_vp, _i64, C.POINTER(_vp), C.POINTER(IpcHandle)]),
    "lrc_peer_buffer_open": (_i32, [_vp, C.POINTER(IpcHandle), C.POINTER(_vp)]),
    "lrc_peer_buffer_close": (_i32, [_vp, _vp]),
    "lrc_peer_buffer_destroy": (_i32, [_vp, _vp]),
    "lrc_set_gather": (_i32, [_vp, C.POINTER(Gather)]),
    "lrc_set_gather_wire": (_i32, [_vp, C.POINTER(GatherWire)]),
    "lrc_gen_rays_single_axis": (_i32, [_vp, _vp, _i64, C.POINTER(SingleAxis), _vp, _vp]),
    "lrc_gen_rays_dual_axis": (_i32, [_vp, _vp, _i64, C.POINTER(DualAxis), C.POINTER(Noise), _vp, _vp, _vp]),
    "lrc_frame_statistics": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "lrc_incident_angles": (_i32, [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "lrc_pack_ply_records": (_i32, [_vp, _vp, _vp, _vp, _vp, C.c_uint32, _i64, _vp, _vp]),
    "lrc_nn_index_build": (_i32, [_vp, _vp, _i64, _dbl, _vp]),
    "lrc_nn_query": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lrc_collision_index_build": (_i32, [_vp, _vp, _i64, _dbl, _vp]),
    "lrc_collision_query": (_i32, [_vp, _vp, _i64, _dbl, C.POINTER(_dbl), _vp, _vp]),
    "lrc_grid_connectivity": (_i32, [_vp, _vp, C.c_int32, _vp, C.c_int32, _vp, _dbl, C.c_int32, _vp, _vp, _vp, _i64, C.POINTER(_i64), _vp]),
    "lrc_astar": (_i32, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_int32, C.POINTER(C.c_int32), C.POINTER(_dbl)]),
    "lrc_set_counting": (_i32, [_vp, _i32]),
    "lrc_counters": (_i32, [_vp, C.POINTER(Counters), _i32, _vp]),
    "lrc_launch_count": (_i64, [_vp]),
    "lrc_set_option": (_i32, [_vp, C.c_char_p, _i64]),
    "lrc_default_l2_persist": (_i32, []),
    "lrc_get_stat": (_i32, [_vp, C.c_char_p, C.POINTER(_i64)]),
    "lrc_kernel_times": (_i32, [_vp, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(C.c_int32)]),
}

_lib = None


def load() -> C.CDLL:
    """dlopen csrc/liblrc.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "This engine has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(ctx, rc: int) -> None:
    if rc != 0:
        msg = load().lrc_last_error(ctx)
        raise LrcError(rc, msg.decode("utf-8", "replace") if msg else "")
