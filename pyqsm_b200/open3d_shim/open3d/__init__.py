"""Stand-in for ``open3d`` covering only what pyQSM's ray-casting path touches
(see README.md next to this package).  Not a fork of Open3D: the scene is
``pyqsm_b200.RaycastingScene``; tensors are torch tensors."""
from . import core, t  # noqa: F401

__version__ = "0.18.0+qsmrt.shim"
