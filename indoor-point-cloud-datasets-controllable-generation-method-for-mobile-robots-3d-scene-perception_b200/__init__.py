"""
B200-native LiDAR ray-casting engine: a drop-in for the ray-cast step of the indoor mobile-LiDAR dataset
generator (reference ``raycast_engine/`` + ``lidar/``; called from ``s3dis_simulator.py:254-288``).

The directory name contains hyphens, so import it through the alias package::

    import lrc_b200 as lrc
    engine = lrc.RaycastEngineGPU()
    points, incident = engine.lidar_intersect_mesh(lrc.create_lidar(intrinsics, pose), mesh)

Sub-packages keep the reference's module names: ``lidar`` (sensor model) and ``raycast_engine`` (engine).
Importing the package needs neither a GPU nor the built library; creating an engine needs both.
"""
from . import lidar, post, raycast_engine, simulator, synthetic, trajectory
from .core import (Context, NoiseConfig, ScanResult, TriangleMesh, get_context, mesh_arrays, pack_labels,
                   rays_per_frame, unpack_labels)
from .lidar import (DualAxisLidar, DualAxisLidarIntrinsics, Indoor8LineLidarIntrinsics, IndoorLidar, LidarIntrinsics,
                    create_lidar, get_lidar_type)
from .raycast_engine import RaycastEngineBase, RaycastEngineGPU
from .simulator import SimFrame, SimRun, room_bounds_of, room_volume, run_simulation
from .post import (LabelTransfer, ScanQuality, SimulationStats, frame_statistics, read_labeled_ply, scan_quality, simulation_stats,
                   write_labeled_ply)
from .trajectory import AutoTrajectoryGenerator, TrajectoryQuality, Waypoint, poses_from_waypoints, shard_range

__version__ = "0.1.0"

__all__ = [
    "lidar", "post", "raycast_engine", "synthetic", "trajectory",
    "ScanQuality", "SimulationStats", "frame_statistics", "scan_quality", "simulation_stats", "write_labeled_ply",
    "read_labeled_ply", "LabelTransfer",
    "Context", "NoiseConfig", "ScanResult", "TriangleMesh", "get_context", "mesh_arrays", "pack_labels",
    "unpack_labels", "rays_per_frame",
    "LidarIntrinsics", "Indoor8LineLidarIntrinsics", "DualAxisLidarIntrinsics", "IndoorLidar", "DualAxisLidar",
    "create_lidar", "get_lidar_type",
    "RaycastEngineBase", "RaycastEngineGPU", "run_simulation", "SimRun", "SimFrame", "room_bounds_of", "room_volume",
    "Waypoint", "poses_from_waypoints", "shard_range", "AutoTrajectoryGenerator", "TrajectoryQuality",
]
