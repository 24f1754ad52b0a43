"""Minimal stand-in for ``open3d.t.geometry.TriangleMesh`` covering exactly the
attribute surface ``pyQSM/viz/ray_casting.py`` touches on a mesh
(``mesh.vertex['positions']``, ``mesh.vertex.positions``,
``mesh.triangle['indices']``, ``.min(dim=0)``, ``.numpy()``, ``get_center()``;
lines :159-160,172-173,245-246,269), for boxes without Open3D."""
from __future__ import annotations

import numpy as np
import torch


class _Attr(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class _T(torch.Tensor):
    """torch tensor whose min/max(dim) return values only, like open3d.core.Tensor."""

    @staticmethod
    def wrap(t):
        return t.as_subclass(_T)

    def min(self, dim=None, **kw):
        r = torch.Tensor.min(self.as_subclass(torch.Tensor)) if dim is None else torch.Tensor.min(self.as_subclass(torch.Tensor), dim=dim).values
        return r

    def max(self, dim=None, **kw):
        r = torch.Tensor.max(self.as_subclass(torch.Tensor)) if dim is None else torch.Tensor.max(self.as_subclass(torch.Tensor), dim=dim).values
        return r


class TriangleMesh:
    def __init__(self, vertex_positions, triangle_indices):
        v = torch.from_numpy(np.ascontiguousarray(vertex_positions, dtype=np.float32))
        t = torch.from_numpy(np.ascontiguousarray(triangle_indices, dtype=np.uint32))
        self.vertex = _Attr(positions=_T.wrap(v))
        self.triangle = _Attr(indices=t)

    def get_center(self):
        return self.vertex["positions"].as_subclass(torch.Tensor).mean(dim=0)

    def get_min_bound(self):
        return self.vertex["positions"].min(dim=0)

    def get_max_bound(self):
        return self.vertex["positions"].max(dim=0)
