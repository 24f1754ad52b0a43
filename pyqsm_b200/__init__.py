"""pyqsm_b200 -- B200-native ray/mesh intersection behind pyQSM's
``RaycastingScene`` call sites (``pyQSM/viz/ray_casting.py``).

Only the hot path lives here: the Open3D-compatible ``RaycastingScene`` (a
ctypes front end of ``libqsmrt.so``, hand-written sm_100a kernels), the
synthetic inputs of the benchmark configurations and the environmental
drivers built on top.  There is no CPU compute path.
"""
from .raycasting_scene import RaycastingScene, INVALID_ID  # noqa: F401



def empty_cache() -> None:
    """Return the device blocks libqsmrt keeps for reuse (freed scenes, commit scratch) to the CUDA driver --
    the counterpart of ``torch.cuda.empty_cache()``.  The cache is bounded by ``QSMRT_CACHE_MB`` (default 1024)."""
    from . import _lib
    _lib.check(_lib.load().qsmrt_release_cached_memory())


__all__ = ["RaycastingScene", "INVALID_ID", "empty_cache"]
