"""``open3d.t.geometry`` names on the ray-casting path: ``RaycastingScene`` (``ray_casting.py:8``) and a
``TriangleMesh`` carrying ``vertex['positions']`` / ``triangle['indices']``."""
from pyqsm_b200.raycasting_scene import RaycastingScene  # noqa: F401
from pyqsm_b200.mesh import TriangleMesh  # noqa: F401
