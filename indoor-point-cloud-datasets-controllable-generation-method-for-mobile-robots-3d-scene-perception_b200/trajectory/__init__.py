"""
Pose format and trajectory generation (reference ``trajectory/`` package: trajectory_generator.py,
auto_trajectory_generator.py).  ``Waypoint.to_pose_matrix()`` is the pose format the ray-cast engine consumes; the
coverage planner produces the waypoints (its occupancy / collision / connectivity passes run on the GPU).
"""
from .trajectory_generator import (TrajectoryQuality, Waypoint, polyline_waypoints, poses_from_waypoints, shard_range)
from .auto_trajectory_generator import AutoTrajectoryGenerator, RoomAnalysis, TrajectoryCandidate
from .collision_detector import CollisionDetector, FurnitureInfo

__all__ = ["Waypoint", "TrajectoryQuality", "poses_from_waypoints", "polyline_waypoints", "shard_range",
           "AutoTrajectoryGenerator", "RoomAnalysis", "TrajectoryCandidate", "CollisionDetector", "FurnitureInfo"]
