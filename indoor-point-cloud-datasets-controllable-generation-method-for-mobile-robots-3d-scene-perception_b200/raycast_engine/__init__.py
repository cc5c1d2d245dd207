"""Ray-cast engines (reference ``raycast_engine/__init__.py``).  Only the GPU engine exists here: the
reference's CPU engine is restated under ``oracle/`` as test infrastructure, never shipped."""
from .base import RaycastEngineBase
from .gpu import RaycastEngineGPU

__all__ = ["RaycastEngineBase", "RaycastEngineGPU"]
