"""ctypes front end of ``libqsmrt_oracle.so`` (the C restatement of Open3D's
CPU ``RaycastingScene``; see ``qsmrt_oracle.c`` for the reference citations).

TEST INFRASTRUCTURE ONLY -- the product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

INVALID_ID = 0xFFFFFFFF
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqsmrt_oracle.so")
_lib = None

BRUTE, BVH = 0, 1


class _Counters(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("tris", C.c_uint64)]


class _Node(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("hi", C.c_float * 3), ("left", C.c_int32), ("right", C.c_int32)]


def build_oracle(force: bool = False) -> str:
    """Compile the oracle with the Makefile next to it (gcc, OpenMP)."""
    src = os.path.join(_HERE, "qsmrt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build_oracle()
    L = C.CDLL(_LIB_PATH)
    vp, fp, u32p, i32p, i64p, u8p = (C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8))
    L.orc_scene_create.restype = vp
    L.orc_scene_destroy.argtypes = [vp]
    L.orc_add_triangles.restype = C.c_uint32
    L.orc_add_triangles.argtypes = [vp, fp, C.c_uint64, u32p, C.c_uint64]
    L.orc_commit.argtypes = [vp]
    L.orc_cast_rays.argtypes = [vp, fp, C.c_uint64, C.c_int, fp, u32p, u32p, fp, fp, C.POINTER(_Counters)]
    L.orc_count_intersections.argtypes = [vp, fp, C.c_uint64, C.c_int, i32p, C.POINTER(_Counters)]
    L.orc_test_occlusions.argtypes = [vp, fp, C.c_uint64, C.c_int, C.c_float, C.c_float, u8p]
    L.orc_list_intersections.argtypes = [vp, fp, C.c_uint64, C.c_int, i64p, i64p, fp, u32p, u32p, fp]
    L.orc_closest_points.argtypes = [vp, fp, C.c_uint64, C.c_int, fp, fp, u32p, u32p, fp, fp]
    L.orc_edge_flags.argtypes = [vp, fp, C.c_uint64, C.c_int, C.c_float, u8p]
    L.orc_num_triangles.restype = C.c_uint64
    L.orc_num_triangles.argtypes = [vp]
    L.orc_sorted_keys.restype = C.POINTER(C.c_uint64)
    L.orc_sorted_keys.argtypes = [vp]
    L.orc_sorted_order.restype = u32p
    L.orc_sorted_order.argtypes = [vp]
    L.orc_nodes.restype = C.POINTER(_Node)
    L.orc_nodes.argtypes = [vp]
    L.orc_box_pad.restype = C.c_float
    L.orc_box_pad.argtypes = [vp]
    L.orc_scene_bounds.argtypes = [vp, fp, fp]
    L.orc_num_threads.restype = C.c_int
    _lib = L
    return L


def _f32(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and (a.ndim < 1 or a.shape[-1] != shape_last):
        raise RuntimeError(f"expected last dim {shape_last}, got shape {a.shape}")
    return a


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class OracleScene:
    """Open3D ``RaycastingScene`` semantics on the CPU.  ``mode`` = BRUTE (every
    triangle; ground truth) or BVH (canonical LBVH; same answers, and the fetch
    counters that define the roofline bytes per ray)."""

    INVALID_ID = INVALID_ID

    def __init__(self):
        self._L = _load()
        self._s = C.c_void_p(self._L.orc_scene_create())
        self.last_counters = (0, 0)

    def __del__(self):
        try:
            if self._s:
                self._L.orc_scene_destroy(self._s)
                self._s = None
        except Exception:
            pass

    @property
    def num_threads(self) -> int:
        return int(self._L.orc_num_threads())

    def add_triangles(self, vertex_positions, triangle_indices) -> int:
        v = _f32(vertex_positions, 3).reshape(-1, 3)
        t = np.ascontiguousarray(triangle_indices, dtype=np.uint32).reshape(-1, 3)
        g = self._L.orc_add_triangles(self._s, _p(v, C.c_float), v.shape[0], _p(t, C.c_uint32), t.shape[0])
        if g == INVALID_ID:
            raise RuntimeError("triangle index out of range")
        return int(g)

    def commit(self):
        self._L.orc_commit(self._s)

    def cast_rays(self, rays, mode=BVH):
        r = _f32(rays, 6)
        shp = r.shape[:-1]
        r2 = r.reshape(-1, 6)
        n = r2.shape[0]
        self.commit()
        out = dict(t_hit=np.empty(n, np.float32), geometry_ids=np.empty(n, np.uint32),
                   primitive_ids=np.empty(n, np.uint32), primitive_uvs=np.empty((n, 2), np.float32),
                   primitive_normals=np.empty((n, 3), np.float32))
        c = _Counters()
        self._L.orc_cast_rays(self._s, _p(r2, C.c_float), n, mode, _p(out["t_hit"], C.c_float),
                              _p(out["geometry_ids"], C.c_uint32), _p(out["primitive_ids"], C.c_uint32),
                              _p(out["primitive_uvs"], C.c_float), _p(out["primitive_normals"], C.c_float),
                              C.byref(c))
        self.last_counters = (int(c.nodes), int(c.tris))
        out["t_hit"] = out["t_hit"].reshape(shp)
        out["geometry_ids"] = out["geometry_ids"].reshape(shp)
        out["primitive_ids"] = out["primitive_ids"].reshape(shp)
        out["primitive_uvs"] = out["primitive_uvs"].reshape(shp + (2,))
        out["primitive_normals"] = out["primitive_normals"].reshape(shp + (3,))
        return out

    def count_intersections(self, rays, mode=BVH):
        r = _f32(rays, 6)
        shp = r.shape[:-1]
        r2 = r.reshape(-1, 6)
        self.commit()
        out = np.empty(r2.shape[0], np.int32)
        c = _Counters()
        self._L.orc_count_intersections(self._s, _p(r2, C.c_float), r2.shape[0], mode, _p(out, C.c_int32), C.byref(c))
        self.last_counters = (int(c.nodes), int(c.tris))
        return out.reshape(shp)

    def test_occlusions(self, rays, tnear=0.0, tfar=float("inf"), mode=BVH):
        r = _f32(rays, 6)
        shp = r.shape[:-1]
        r2 = r.reshape(-1, 6)
        self.commit()
        out = np.empty(r2.shape[0], np.uint8)
        self._L.orc_test_occlusions(self._s, _p(r2, C.c_float), r2.shape[0], mode, tnear, tfar, _p(out, C.c_uint8))
        return out.astype(bool).reshape(shp)

    def list_intersections(self, rays, mode=BVH):
        r = _f32(rays, 6)
        r2 = r.reshape(-1, 6)
        n = r2.shape[0]
        counts = self.count_intersections(r2, mode).astype(np.int64)
        splits = np.zeros(n + 1, np.int64)
        np.cumsum(counts, out=splits[1:])
        k = int(splits[-1])
        out = dict(ray_splits=splits, ray_ids=np.empty(k, np.int64), t_hit=np.empty(k, np.float32),
                   geometry_ids=np.empty(k, np.uint32), primitive_ids=np.empty(k, np.uint32),
                   primitive_uvs=np.empty((k, 2), np.float32))
        self._L.orc_list_intersections(self._s, _p(r2, C.c_float), n, mode, _p(splits, C.c_int64),
                                       _p(out["ray_ids"], C.c_int64), _p(out["t_hit"], C.c_float),
                                       _p(out["geometry_ids"], C.c_uint32), _p(out["primitive_ids"], C.c_uint32),
                                       _p(out["primitive_uvs"], C.c_float))
        return out

    def compute_closest_points(self, query_points, mode=BVH):
        """Open3D ComputeClosestPoints keys (points, geometry_ids, primitive_ids, primitive_uvs,
        primitive_normals) plus ``distance``."""
        p = _f32(query_points, 3)
        shp = p.shape[:-1]
        p2 = p.reshape(-1, 3)
        n = p2.shape[0]
        self.commit()
        out = dict(points=np.empty((n, 3), np.float32), distance=np.empty(n, np.float32),
                   geometry_ids=np.empty(n, np.uint32), primitive_ids=np.empty(n, np.uint32),
                   primitive_uvs=np.empty((n, 2), np.float32), primitive_normals=np.empty((n, 3), np.float32))
        self._L.orc_closest_points(self._s, _p(p2, C.c_float), n, mode, _p(out["points"], C.c_float), _p(out["distance"], C.c_float),
                                   _p(out["geometry_ids"], C.c_uint32), _p(out["primitive_ids"], C.c_uint32),
                                   _p(out["primitive_uvs"], C.c_float), _p(out["primitive_normals"], C.c_float))
        for k in out:
            out[k] = out[k].reshape(shp + out[k].shape[1:])
        return out

    def compute_distance(self, query_points, mode=BVH):
        return self.compute_closest_points(query_points, mode)["distance"]

    def compute_occupancy(self, query_points, mode=BVH):
        p = _f32(query_points, 3)
        rays = np.concatenate([p, np.ones_like(p)], axis=-1)
        return (self.count_intersections(rays, mode) % 2 == 1).astype(np.float32)

    def compute_signed_distance(self, query_points, mode=BVH):
        d = self.compute_distance(query_points, mode)
        return np.where(self.compute_occupancy(query_points, mode) > 0, -d, d).astype(np.float32)

    def edge_flags(self, rays, eps=1e-6, mode=BVH):
        """bit0: a triangle plane is crossed with a barycentric within ``eps`` of
        an edge; bit1: two accepted hits share the same t (tie)."""
        r2 = _f32(rays, 6).reshape(-1, 6)
        self.commit()
        out = np.empty(r2.shape[0], np.uint8)
        self._L.orc_edge_flags(self._s, _p(r2, C.c_float), r2.shape[0], mode, eps, _p(out, C.c_uint8))
        return out

    # ---- builder introspection (for the LBVH builder parity tests)
    def sorted_keys(self):
        self.commit()
        n = int(self._L.orc_num_triangles(self._s))
        return np.ctypeslib.as_array(self._L.orc_sorted_keys(self._s), shape=(n,)).copy() if n else np.empty(0, np.uint64)

    def sorted_order(self):
        self.commit()
        n = int(self._L.orc_num_triangles(self._s))
        return np.ctypeslib.as_array(self._L.orc_sorted_order(self._s), shape=(n,)).copy() if n else np.empty(0, np.uint32)

    def nodes(self):
        """(lo[n-1,3], hi[n-1,3], left[n-1], right[n-1]) of the canonical LBVH."""
        self.commit()
        n = int(self._L.orc_num_triangles(self._s))
        if n < 2:
            z = np.empty((0, 3), np.float32)
            return z, z, np.empty(0, np.int32), np.empty(0, np.int32)
        raw = np.ctypeslib.as_array(C.cast(self._L.orc_nodes(self._s), C.POINTER(C.c_uint8)), shape=((n - 1) * 32,))
        f = raw.view(np.float32).reshape(n - 1, 8)
        i = raw.view(np.int32).reshape(n - 1, 8)
        return f[:, 0:3].copy(), f[:, 3:6].copy(), i[:, 6].copy(), i[:, 7].copy()

    def box_pad(self) -> float:
        self.commit()
        return float(self._L.orc_box_pad(self._s))

    def scene_bounds(self):
        self.commit()
        lo = np.empty(3, np.float32)
        hi = np.empty(3, np.float32)
        self._L.orc_scene_bounds(self._s, _p(lo, C.c_float), _p(hi, C.c_float))
        return lo, hi
