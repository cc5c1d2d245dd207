"""Abstract engine interface (reference ``raycast_engine/raycast_engine.py:16-62``)."""
from abc import ABC, abstractmethod

import numpy as np


class RaycastEngineBase(ABC):
    """Mesh/ray intersection engine.

    The reference notes that its API assumes "a scene is only used for raycasting once"
    (raycast_engine.py:20-23) and rebuilds the acceleration structure on every call.  The GPU engine keeps
    the same two-method interface but caches the BVH between calls on an unchanged mesh.
    """

    @abstractmethod
    def __init__(self):
        pass

    @abstractmethod
    def rays_intersect_mesh(self, rays: np.ndarray, mesh):
        """rays (N,6) float32, mesh: TriangleMesh-like -> hit points (M,3) float32, M <= N, ray order."""

    @abstractmethod
    def lidar_intersect_mesh(self, lidar, mesh):
        """lidar: posed sensor, mesh: TriangleMesh-like -> (points (M,3) float32, incident_angles (M,) float64)."""
