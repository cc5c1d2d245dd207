"""
What the reference does with a frame right after the ray-cast call -- kept on the GPU:

* ``ScanQuality`` per frame                        reference s3dis_simulator.py:276-284, containers/s3dis_sim_frame.py:11-40
* ``SimulationStats`` over the trajectory          reference containers/s3dis_sim_scene.py:29-55,157-179,228-247
* the labelled PLY (8 attributes, 19 B per vertex)  reference containers/s3dis_sim_scene.py:614-641 (writer),
                                                    lidar_net_bbox_visualizer.py:72-126 (reader)

The sums and the record packing run in ``csrc/post.cu`` (``lrc_frame_statistics``, ``lrc_pack_ply_records``); only
P small records / the finished byte stream cross PCIe.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as nat
from .core import Context, ScanResult, _ptr

PLY_RECORD_BYTES = 19
PLY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1"),
                      ("sem", "<u2"), ("ins", "<u2")])
assert PLY_DTYPE.itemsize == PLY_RECORD_BYTES
DEFAULT_GREY = 0x7F7F7F        # (0.5 * 255).astype(uint8) per channel: reference s3dis_sim_scene.py:584-594 + :483


@dataclass
class ScanQuality:
    """Same fields as the reference's ``ScanQuality`` (containers/s3dis_sim_frame.py:11-40)."""
    coverage_ratio: float
    num_points: int
    incident_angle_mean: float
    incident_angle_std: float
    scan_density: float
    range_mean: float
    range_std: float

    def to_dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)


@dataclass
class SimulationStats:
    """Same fields as the reference's ``SimulationStats`` (containers/s3dis_sim_scene.py:29-55)."""
    total_frames: int
    total_points: int
    average_coverage: float
    average_scan_density: float
    average_incident_angle: float
    average_range: float
    simulation_time: float
    frames_per_second: float

    def to_dict(self) -> Dict[str, Any]:
        return dict(self.__dict__)


_STATS_DTYPE = np.dtype([("num_points", "<i8"), ("incident_mean", "<f8"), ("incident_std", "<f8"),
                         ("range_mean", "<f8"), ("range_std", "<f8")])
assert _STATS_DTYPE.itemsize == C.sizeof(nat.FrameStats)


def frame_statistics(ctx: Context, result: ScanResult) -> np.ndarray:
    """Per-frame sums of a scan on the GPU -> structured array (P,) with num_points, incident_mean/std, range_mean/std."""
    P = result.num_frames
    out = torch.empty((max(P, 1), _STATS_DTYPE.itemsize), dtype=torch.uint8, device=ctx.device)
    with torch.cuda.device(ctx.device):
        inc = result.incident if result.incident.numel() == result.points.shape[0] and result.points.shape[0] > 0 else None
        nat.check(ctx._h, ctx._lib.lrc_frame_statistics(ctx._h, _ptr(result.points), _ptr(inc), _ptr(result.frame_offset), P,
                                                        _ptr(out), ctx._stream()))
    return out[:P].cpu().numpy().view(_STATS_DTYPE).reshape(-1).copy()


def scan_quality(ctx: Context, result: ScanResult, total_points_per_scan: int, room_volume: float) -> List[ScanQuality]:
    """The ``ScanQuality`` the reference's frame loop builds for every frame (s3dis_simulator.py:276-284)."""
    st = frame_statistics(ctx, result)
    return [ScanQuality(coverage_ratio=int(s["num_points"]) / total_points_per_scan, num_points=int(s["num_points"]),
                        incident_angle_mean=float(s["incident_mean"]), incident_angle_std=float(s["incident_std"]),
                        scan_density=int(s["num_points"]) / room_volume, range_mean=float(s["range_mean"]),
                        range_std=float(s["range_std"])) for s in st]


def simulation_stats(qualities: Sequence[ScanQuality], simulation_time: float) -> SimulationStats:
    """``S3DISSimScene.compute_statistics`` (s3dis_sim_scene.py:228-247).  The reference's exporter recomputes this
    with ``simulation_time=0.0`` (:254) and therefore always writes 0 FPS; here the measured time is kept."""
    if not qualities:
        return SimulationStats(0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    return SimulationStats(
        total_frames=len(qualities), total_points=int(sum(q.num_points for q in qualities)),
        average_coverage=float(np.mean([q.coverage_ratio for q in qualities])),
        average_scan_density=float(np.mean([q.scan_density for q in qualities])),
        average_incident_angle=float(np.mean([q.incident_angle_mean for q in qualities])),
        average_range=float(np.mean([q.range_mean for q in qualities])),
        simulation_time=float(simulation_time),
        frames_per_second=len(qualities) / simulation_time if simulation_time > 0 else 0.0)


def ply_header(num_vertices: int) -> bytes:
    """The 12 header lines of reference s3dis_sim_scene.py:621-632, byte for byte."""
    return (b"ply\nformat binary_little_endian 1.0\n" + b"element vertex %d\n" % num_vertices +
            b"property float x\nproperty float y\nproperty float z\n"
            b"property uchar red\nproperty uchar green\nproperty uchar blue\n"
            b"property ushort sem\nproperty ushort ins\nend_header\n")


def pack_rgb(rgb) -> np.ndarray:
    """(T,3) uint8 colours -> uint32 red | green << 8 | blue << 16 (the ``tri_rgb`` table of lrc_pack_ply_records)."""
    c = np.asarray(rgb, dtype=np.uint32).reshape(-1, 3)
    return (c[:, 0] & 255) | ((c[:, 1] & 255) << 8) | ((c[:, 2] & 255) << 16)


def pack_ply_records(ctx: Context, result: ScanResult, tri_rgb: Optional[torch.Tensor] = None,
                     default_rgb: int = DEFAULT_GREY) -> torch.Tensor:
    """Device byte tensor (19 * M,) holding the vertex records of the whole cloud, frames concatenated in order
    (what ``_export_combined_pointcloud_with_labels`` stacks, s3dis_sim_scene.py:336-377)."""
    M = int(result.points.shape[0])
    out = torch.empty(max(M * PLY_RECORD_BYTES, 16), dtype=torch.uint8, device=ctx.device)
    with torch.cuda.device(ctx.device):
        lab = result.label if result.label.numel() == M and M > 0 else None
        prim = result.prim_id if (tri_rgb is not None and result.prim_id.numel() == M and M > 0) else None
        if tri_rgb is not None and prim is None and M > 0:
            raise ValueError("per-triangle colours need the scan's prim_id array")
        nat.check(ctx._h, ctx._lib.lrc_pack_ply_records(ctx._h, _ptr(result.points), _ptr(lab), _ptr(prim),
                                                        _ptr(tri_rgb) if prim is not None else None,
                                                        int(default_rgb) & 0xFFFFFF, M, _ptr(out), ctx._stream()))
    return out[:M * PLY_RECORD_BYTES]


def write_labeled_ply(ctx: Context, path, result: ScanResult, tri_rgb=None, default_rgb: int = DEFAULT_GREY) -> int:
    """``combined_pointcloud_with_label.ply`` for a scan: header + records packed on the GPU.  Returns bytes written.
    Byte-identical to the reference writer for the same points / colours / labels."""
    if tri_rgb is not None and not isinstance(tri_rgb, torch.Tensor):
        tri_rgb = torch.from_numpy(np.ascontiguousarray(tri_rgb, dtype=np.uint32).view(np.int32)).to(ctx.device)
    rec = pack_ply_records(ctx, result, tri_rgb, default_rgb)
    host = torch.empty(rec.numel(), dtype=torch.uint8).pin_memory() if rec.numel() else torch.empty(0, dtype=torch.uint8)
    host.copy_(rec)
    hdr = ply_header(int(result.points.shape[0]))
    with open(path, "wb") as f:
        f.write(hdr)
        host.numpy().tofile(f)
    return len(hdr) + int(host.numel())


class LabelTransfer:
    """Exact 1-nearest-neighbour transfer of colours / semantic / instance labels from annotated points to hit points --
    what ``S3DISSimScene._get_colors_and_labels_from_s3dis`` does per frame with a scikit-learn ball tree
    (reference containers/s3dis_sim_scene.py:413-424), here on the GPU (``lrc_nn_index_build`` / ``lrc_nn_query``).

    ``points``: (n,3) annotated points (float64, as loaded from the S3DIS annotation files); ``colors``: (n,3) floats
    in [0,1] (reference convention, :483 turns them into bytes with ``(c * 255).astype(uint8)``) or uint8;
    ``semantic`` / ``instance``: (n,) integer labels.  Indices agree with the ball tree bit for bit except at exact
    distance ties (smaller index here)."""

    def __init__(self, ctx: Context, points, colors=None, semantic=None, instance=None, cell: float = 0.0):
        self.ctx = ctx
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.n = len(pts)
        self._cell = float(cell)
        dev = ctx.device
        with torch.cuda.device(dev):
            self._pts = torch.from_numpy(pts).to(dev)
            self._build_index()
        self._lab = self._rgb = None
        if semantic is not None or instance is not None:
            sem = np.zeros(self.n, np.uint32) if semantic is None else np.asarray(semantic).astype(np.uint32)
            ins = np.zeros(self.n, np.uint32) if instance is None else np.asarray(instance).astype(np.uint32)
            lab = (sem & 0xFFFF) | ((ins & 0xFFFF) << 16)
            self._lab = torch.from_numpy(lab.view(np.int32)).to(dev)
        if colors is not None:
            c = np.asarray(colors)
            if c.dtype != np.uint8:
                c = (c * 255).astype(np.uint8)                   # reference :483
            self._rgb = torch.from_numpy(pack_rgb(c).view(np.int32)).to(dev)

    def _build_index(self) -> None:
        ctx = self.ctx
        nat.check(ctx._h, ctx._lib.lrc_nn_index_build(ctx._h, _ptr(self._pts), self.n, self._cell, ctx._stream()))
        # the context holds ONE nearest-neighbour index; the label / colour tables live here.  Remember which build is
        # ours: if another LabelTransfer (another room) has re-targeted the slot, query() re-bins OUR points first --
        # the reference fits a fresh tree per call (s3dis_sim_scene.py:413-424), so objects never share state there.
        self._generation = ctx.stat("nn_generation")

    def query(self, points, want_distance: bool = False) -> Dict[str, torch.Tensor]:
        """points: (M,3) float32 (device tensor or ndarray).  -> device tensors: index int32, label / rgb int32 holding
        uint32 bits (present when the tables were given), distance float64 (optional)."""
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            if ctx.stat("nn_generation") != self._generation:
                self._build_index()
            q = points if isinstance(points, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32))
            q = q.to(device=ctx.device, dtype=torch.float32).contiguous().reshape(-1, 3)
            M = q.shape[0]
            idx = torch.empty(max(M, 1), dtype=torch.int32, device=ctx.device)[:M]
            dist = torch.empty(max(M, 1), dtype=torch.float64, device=ctx.device)[:M] if want_distance else None
            lab = torch.empty(max(M, 1), dtype=torch.int32, device=ctx.device)[:M] if self._lab is not None else None
            rgb = torch.empty(max(M, 1), dtype=torch.int32, device=ctx.device)[:M] if self._rgb is not None else None
            nat.check(ctx._h, ctx._lib.lrc_nn_query(ctx._h, _ptr(q), M, _ptr(idx), _ptr(dist), _ptr(self._lab), _ptr(lab),
                                                    _ptr(self._rgb), _ptr(rgb), ctx._stream()))
        out = {"index": idx}
        if lab is not None:
            out["label"] = lab
        if rgb is not None:
            out["rgb"] = rgb
        if dist is not None:
            out["distance"] = dist
        return out

    def relabel(self, result: ScanResult) -> ScanResult:
        """A copy of ``result`` whose ``label`` comes from the annotated points and whose ``prim_id`` holds the index
        of the nearest annotated point (so ``write_labeled_ply(..., tri_rgb=self.rgb_table)`` colours by neighbour)."""
        r = self.query(result.points)
        return ScanResult(points=result.points, incident=result.incident, prim_id=r["index"],
                          label=r.get("label", result.label), ray_idx=result.ray_idx, frame_offset=result.frame_offset,
                          frame_offset_host=result.frame_offset_host)

    @property
    def rgb_table(self) -> Optional[torch.Tensor]:
        return self._rgb


def read_labeled_ply(path) -> Dict[str, np.ndarray]:
    """Parse the 8-attribute binary PLY the way the reference's consumer does (lidar_net_bbox_visualizer.py:72-126:
    header lines until ``end_header``, vertex count from ``element vertex``, 12 + 3 bytes skipped, ``HH`` labels)."""
    with open(path, "rb") as f:
        lines = []
        while True:
            line = f.readline().decode("utf-8").strip()
            lines.append(line)
            if line == "end_header":
                break
        props = [ln for ln in lines if ln.startswith("property")]
        if not (any("sem" in ln for ln in props) and any("ins" in ln for ln in props)):
            raise ValueError("semantic or instance attributes missing (expected x,y,z,r,g,b,sem,ins)")
        n = 0
        for ln in lines:
            if ln.startswith("element vertex"):
                n = int(ln.split()[-1])
        rec = np.fromfile(f, dtype=PLY_DTYPE, count=n)
    return {"points": np.stack([rec["x"], rec["y"], rec["z"]], axis=1),
            "colors": np.stack([rec["red"], rec["green"], rec["blue"]], axis=1),
            "semantic_labels": rec["sem"].copy(), "instance_labels": rec["ins"].copy()}
