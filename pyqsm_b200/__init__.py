"""pyqsm_b200 -- B200-native ray/mesh intersection behind pyQSM's
``RaycastingScene`` call sites (``pyQSM/viz/ray_casting.py``).

Only the hot path lives here: the Open3D-compatible ``RaycastingScene`` (a
ctypes front end of ``libqsmrt.so``, hand-written sm_100a kernels), the
synthetic inputs of the benchmark configurations and the environmental
drivers built on top.  There is no CPU compute path.
"""
from .raycasting_scene import RaycastingScene, INVALID_ID  # noqa: F401

__all__ = ["RaycastingScene", "INVALID_ID"]
