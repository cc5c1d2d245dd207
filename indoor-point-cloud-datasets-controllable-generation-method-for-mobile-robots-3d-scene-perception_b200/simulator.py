"""
The simulator's frame loop as one batched GPU pass -- what ``S3DISSimulator.run_simulation(waypoints)`` does per room
(reference s3dis_simulator.py:220-296; the README calls it ``simulate_room``, :174).

Reference, per waypoint: pose matrix -> ``create_lidar`` -> ``raycast_engine.lidar_intersect_mesh`` -> ``ScanQuality``
-> ``S3DISSimFrame`` -> (after the loop) ``compute_statistics``.  Here the whole trajectory goes through
``RaycastEngineGPU.simulate`` (one launch sequence), the ScanQuality sums are reduced on the GPU (``lrc_frame_statistics``)
and the frames come back as numpy views of one transfer.  Engine errors are raised, not swallowed (the reference's
blanket ``except Exception`` at :271-273 turns them into empty frames).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .core import NoiseConfig, ScanResult
from .post import ScanQuality, SimulationStats, scan_quality, simulation_stats
from .trajectory import Waypoint, poses_from_waypoints


def room_bounds_of(mesh) -> Dict[str, float]:
    """Bounds dict of the mesh vertices, as ``S3DISSimulator.load_scene`` computes it (s3dis_simulator.py:96-104)."""
    v = np.asarray(mesh.vertices if hasattr(mesh, "vertices") else mesh[0])
    return {"x_min": float(v[:, 0].min()), "x_max": float(v[:, 0].max()), "y_min": float(v[:, 1].min()),
            "y_max": float(v[:, 1].max()), "z_min": float(v[:, 2].min()), "z_max": float(v[:, 2].max())}


def room_volume(bounds: Dict[str, float]) -> float:
    """``RoomBounds.get_volume`` (containers/s3dis_scene.py): product of the three extents."""
    return (bounds["x_max"] - bounds["x_min"]) * (bounds["y_max"] - bounds["y_min"]) * (bounds["z_max"] - bounds["z_min"])


@dataclass
class SimFrame:
    """The per-frame record of the reference (``S3DISSimFrame``, containers/s3dis_sim_frame.py:90-101) plus the
    per-point labels and triangle ids this engine produces."""
    frame_index: int
    points: np.ndarray               # (m,3) float32
    incident_angles: np.ndarray      # (m,)  float64
    scan_quality: ScanQuality
    labels: Optional[np.ndarray] = None      # (m,) uint32 = sem | ins << 16
    prim_id: Optional[np.ndarray] = None     # (m,) uint32

    def get_num_points(self) -> int:
        return len(self.points)


@dataclass
class SimRun:
    frames: List[SimFrame]
    statistics: SimulationStats
    scan: ScanResult                      # device-resident result (for write_labeled_ply, LabelTransfer ...)
    simulation_time: float = 0.0
    config: Dict[str, Any] = field(default_factory=dict)

    def get_total_points(self) -> int:
        return sum(f.get_num_points() for f in self.frames)


def run_simulation(engine, waypoints: Sequence[Waypoint], lidar_config, mesh, noise: Optional[NoiseConfig] = None,
                   bounds: Optional[Dict[str, float]] = None) -> SimRun:
    """== ``S3DISSimulator.run_simulation`` (s3dis_simulator.py:220-296) for one room.

    ``noise=None`` reproduces the bit-exact, noise-free path; pass ``NoiseConfig.from_intrinsics(lidar_config, seed)``
    for the dual-axis sensor's angle noise and dropout (reference indoor_lidar.py:270-272,292-294)."""
    bounds = bounds or room_bounds_of(mesh)
    total_points_per_scan = lidar_config.get_total_points_per_scan()            # :250
    volume = room_volume(bounds)                                                 # :251
    start = time.time()                                                          # :247
    poses = poses_from_waypoints(waypoints)                                      # :256
    if len(poses) == 0:
        empty = engine.simulate(np.zeros((0, 4, 4)), lidar_config, mesh, noise)
        return SimRun([], simulation_stats([], 0.0), empty, 0.0)
    scan = engine.simulate(poses, lidar_config, mesh, noise)                     # :257-263, all frames at once
    qualities = scan_quality(engine.ctx, scan, total_points_per_scan, volume)   # :276-284 on the GPU
    host = scan.numpy()
    off = host["frame_offset"]
    frames = []
    for i in range(len(poses)):                                                  # :287-288
        a, b = int(off[i]), int(off[i + 1])
        pts = host["points"][a:b]
        inc = host["incident"][a:b] if b > a else np.empty(0)                    # raycast_engine_cpu.py:109
        frames.append(SimFrame(i, pts, inc, qualities[i], host["label"][a:b], host["prim_id"][a:b]))
    simulation_time = time.time() - start                                        # :291
    return SimRun(frames, simulation_stats(qualities, simulation_time), scan, simulation_time,
                  {"total_points_per_scan": total_points_per_scan, "room_volume": volume, "bounds": bounds})
