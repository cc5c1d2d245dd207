"""Sensor model (reference ``lidar/__init__.py:6-16``: same six public names)."""
from .intrinsics import LidarIntrinsics, Indoor8LineLidarIntrinsics, DualAxisLidarIntrinsics
from .sensors import IndoorLidar, DualAxisLidar, create_lidar, get_lidar_type

__all__ = [
    "LidarIntrinsics",
    "Indoor8LineLidarIntrinsics",
    "DualAxisLidarIntrinsics",
    "IndoorLidar",
    "DualAxisLidar",
    "create_lidar",
    "get_lidar_type",
]
