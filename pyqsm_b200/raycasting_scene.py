"""``RaycastingScene`` -- drop-in for ``open3d.t.geometry.RaycastingScene`` on the
ray-casting path of wischmcj/pyQSM (``pyQSM/viz/ray_casting.py``).

The reference binds the class at import (``ray_casting.py:8``,
``from open3d.t.geometry import RaycastingScene as rcs``) and then uses
``rcs()`` (:65,155,218,241,275,316), ``add_triangles`` (:66,156,219,242,276,
317), ``create_rays_pinhole`` (:222,230,277,318), ``cast_rays`` (:223,231,279,
319), ``list_intersections`` (:168) and ``compute_occupancy`` (:69).  Same
names, argument meaning, result keys/dtypes/shapes and ``RuntimeError``s here;
the arithmetic runs in ``libqsmrt.so`` (hand-written sm_100a kernels) through
its C ABI with torch tensors as the zero-copy ray and hit buffers.

Results are ``torch`` tensors: on the CPU by default (so ``.numpy()``,
``.isfinite()``, mask indexing and ``.reshape`` behave as the reference
expects), or left on the GPU with ``output_device='cuda'``.

``cast_rays`` returns all five Open3D keys, but only the two the reference
reads (``t_hit``, ``primitive_ids``; ``ray_casting.py:280-289,320-322``) cross
PCIe eagerly -- the other three stay on the GPU and are copied on first access
(``CastResult``), 8 instead of 32 bytes per ray.  ``outputs="all"`` copies
everything at once, ``outputs=("t_hit", ...)`` computes only the named keys.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib

INVALID_ID = 0xFFFFFFFF


def _unwrap(x):
    """numpy / torch / Open3D tensor / list -> numpy or torch, dtype preserved."""
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x
    if hasattr(x, "numpy") and callable(x.numpy):          # open3d.core.Tensor
        return x.numpy()
    return None


def _to_torch(x, want: torch.dtype, name: str, device: torch.device, strict=True) -> torch.Tensor:
    y = _unwrap(x)
    if y is None:                                           # python list / scalar: no dtype to check
        y = np.asarray(x, dtype={torch.float32: np.float32, torch.uint32: np.uint32}[want])
    if isinstance(y, np.ndarray):
        if strict and y.dtype != {torch.float32: np.float32, torch.uint32: np.uint32}[want]:
            raise RuntimeError(f"{name} has dtype {y.dtype}, but it must be {str(want).replace('torch.', '')}")
        y = torch.from_numpy(np.ascontiguousarray(y))
    else:
        if strict and y.dtype != want:
            raise RuntimeError(f"{name} has dtype {y.dtype}, but it must be {want}")
    return y.to(device=device, dtype=want).contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)


def _mesh_arrays(mesh):
    """Duck-typed open3d.t.geometry.TriangleMesh: vertex['positions'] / triangle['indices']."""
    def get(container, key):
        if hasattr(container, key):
            return getattr(container, key)
        return container[key]
    return get(mesh.vertex, "positions"), get(mesh.triangle, "indices")


def _to_host(t: torch.Tensor) -> torch.Tensor:
    """Device tensor -> CPU tensor.  Larger results land in page-locked memory (torch's caching host allocator keeps
    the blocks), which roughly doubles the PCIe rate of this copy and of a later copy back to the device --
    ``create_rays_pinhole`` -> ``cast_rays`` is the reference's own sequence (``ray_casting.py:222-223``)."""
    if t.device.type != "cuda" or t.numel() * t.element_size() < (1 << 18):
        return t.cpu()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t)
    return host


CAST_KEYS = ("t_hit", "geometry_ids", "primitive_ids", "primitive_uvs", "primitive_normals")
_CAST_DTYPE = {"t_hit": torch.float32, "geometry_ids": torch.uint32, "primitive_ids": torch.uint32,
               "primitive_uvs": torch.float32, "primitive_normals": torch.float32}
_CAST_TAIL = {"t_hit": (), "geometry_ids": (), "primitive_ids": (), "primitive_uvs": (2,), "primitive_normals": (3,)}
EAGER_KEYS = ("t_hit", "primitive_ids")       # what the reference consumes (ray_casting.py:280-289,320-322)


class CastResult(dict):
    """``cast_rays`` result: a dict with Open3D's five keys.  Entries that were
    left on the GPU are copied to the host the first time they are read; until
    then they cost no PCIe traffic.  Behaves like a plain dict otherwise."""

    def __init__(self, eager: dict, pending: dict):
        super().__init__(eager)
        self._pending = dict(pending)          # key -> device tensor (already shaped)

    def _fetch(self, key):
        t = _to_host(self._pending.pop(key))
        dict.__setitem__(self, key, t)
        return t

    def __missing__(self, key):
        if key in self._pending:
            return self._fetch(key)
        raise KeyError(key)

    def _fetch_all(self):
        for k in list(self._pending):
            self._fetch(k)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._pending

    def __iter__(self):
        return iter([k for k in CAST_KEYS if k in self] + [k for k in dict.keys(self) if k not in CAST_KEYS])

    def __len__(self):
        return dict.__len__(self) + len(self._pending)

    def keys(self):
        return list(iter(self))

    def items(self):
        self._fetch_all()
        return dict.items(self)

    def values(self):
        self._fetch_all()
        return dict.values(self)

    def pending(self):
        """Keys still resident on the GPU (diagnostics / tests)."""
        return tuple(self._pending)


def _occupancy_directions(nsamples: int):
    """(1,1,1) first (the single-sample direction), then a fixed seeded set of unit vectors."""
    dirs = [(1.0, 1.0, 1.0)]
    if nsamples > 1:
        rng = np.random.default_rng(42)
        v = rng.uniform(-1.0, 1.0, size=(nsamples - 1, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        dirs += [tuple(float(x) for x in r) for r in v]
    return dirs


class RaycastingScene:
    """Open3D-compatible ray casting scene running on one B200.

    ``nthreads`` is accepted and ignored (Open3D: Embree threads).  ``device``
    is the CUDA device (``None`` = torch's current CUDA device; Open3D-style
    ``'CPU:0'`` is accepted and means "return CPU tensors").  There is no CPU
    compute path.
    """

    INVALID_ID = INVALID_ID

    def __init__(self, nthreads: int = 0, device=None, output_device=None):
        self._init_common(device, output_device)
        h = C.c_void_p()
        _lib.check(self._L.qsmrt_scene_create(self.device.index, C.byref(h)))
        self._h = h

    def _init_common(self, device, output_device):
        self._h = None
        self._L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("pyqsm_b200.RaycastingScene needs a CUDA device (B200); there is no CPU fallback")
        dev = None
        if device is not None and not (isinstance(device, str) and device.upper().startswith("CPU")):
            s = str(device).lower().replace("cuda:", "")
            dev = int(s) if s.isdigit() else torch.device(device).index
        if dev is None:
            dev = torch.cuda.current_device()
        self.device = torch.device("cuda", dev)
        self.output_device = torch.device(output_device) if output_device is not None else torch.device("cpu")
        if self.output_device.type == "cuda":
            self.output_device = self.device

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._L.qsmrt_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------- geometry
    def add_triangles(self, vertex_positions, triangle_indices=None) -> int:
        """``add_triangles(mesh)`` or ``add_triangles(vertex_positions, triangle_indices)``;
        returns the geometry id (0, 1, ...).  The mesh is copied."""
        if triangle_indices is None:
            # mesh overload: Open3D casts (GetVertexPositions().To(Float32), GetTriangleIndices().To(UInt32)) -- tensor
            # meshes from from_legacy / create_cylinder carry Int64 indices -- only the two-tensor overload is strict
            vertex_positions, triangle_indices = _mesh_arrays(vertex_positions)
            idx = _unwrap(triangle_indices)
            if idx is None:
                idx = np.asarray(triangle_indices)
            if isinstance(idx, np.ndarray):
                idx = torch.from_numpy(np.ascontiguousarray(idx if idx.dtype != np.uint64 else idx.astype(np.int64)))
            if idx.dtype not in (torch.uint32,) and idx.numel():
                wide = idx.to(torch.int64)
                if int(wide.min()) < 0 or int(wide.max()) > 0xFFFFFFFF:
                    raise RuntimeError("triangle_indices do not fit in UInt32")
            v = _to_torch(vertex_positions, torch.float32, "vertex_positions", self.device, strict=False)
            t = _to_torch(idx, torch.uint32, "triangle_indices", self.device, strict=False)
        else:
            v = _to_torch(vertex_positions, torch.float32, "vertex_positions", self.device)
            t = _to_torch(triangle_indices, torch.uint32, "triangle_indices", self.device)
        if v.ndim != 2 or v.shape[1] != 3:
            raise RuntimeError(f"vertex_positions has shape {tuple(v.shape)}, but it must be (N, 3)")
        if t.ndim != 2 or t.shape[1] != 3:
            raise RuntimeError(f"triangle_indices has shape {tuple(t.shape)}, but it must be (N, 3)")
        gid = C.c_uint32()
        # the ABI copies on the legacy default stream: tensors produced on torch's current (possibly non-blocking)
        # stream must be complete first
        torch.cuda.current_stream(self.device).synchronize()
        _lib.check(self._L.qsmrt_add_triangles(self._h, _ptr(v), v.shape[0], _ptr(t), t.shape[0], 1, C.byref(gid)))
        return int(gid.value)

    def add_cylinders(self, centers, axes, radii, heights, resolution: int = 20, split: int = 4) -> int:
        """Register a cylinder QSM (``cyl_details`` of ``pyQSM/qsm_generation.py:171-178``: centre, axis,
        radius, height per cylinder) as one geometry; the closed cylinder meshes (Open3D ``create_cylinder``
        topology) are generated on the GPU.  Returns the geometry id."""
        rec = torch.cat([torch.as_tensor(np.asarray(centers), dtype=torch.float32).reshape(-1, 3),
                         torch.as_tensor(np.asarray(axes), dtype=torch.float32).reshape(-1, 3),
                         torch.as_tensor(np.asarray(radii), dtype=torch.float32).reshape(-1, 1),
                         torch.as_tensor(np.asarray(heights), dtype=torch.float32).reshape(-1, 1)], dim=1)
        rec = rec.to(self.device).contiguous()
        gid = C.c_uint32()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream(self.device).synchronize()
            _lib.check(self._L.qsmrt_add_cylinders(self._h, _ptr(rec), rec.shape[0], int(resolution), int(split), 1, C.byref(gid)))
        return int(gid.value)

    def geometry(self, geometry_id: int):
        """(vertex_positions float32 [V,3], triangle_indices uint32 [T,3]) of a registered geometry."""
        nv, nt = C.c_uint64(), C.c_uint64()
        _lib.check(self._L.qsmrt_geometry_size(self._h, int(geometry_id), C.byref(nv), C.byref(nt)))
        with torch.cuda.device(self.device):
            v = torch.empty(nv.value, 3, dtype=torch.float32, device=self.device)
            t = torch.empty(nt.value, 3, dtype=torch.uint32, device=self.device)
            _lib.check(self._L.qsmrt_copy_geometry(self._h, int(geometry_id), _ptr(v), _ptr(t), self._stream()))
            return self._out(v), self._out(t)

    def commit(self) -> float:
        """Build the LBVH now (queries do it lazily); returns the build time in ms."""
        ms = C.c_float()
        _lib.check(self._L.qsmrt_commit(self._h, self._stream(), C.byref(ms)))
        return float(ms.value)

    def set_option(self, name: str, value) -> None:
        """Per-scene tuning / test hook (``enum qsmrt_option`` in ``include/qsmrt.h``): builder options take effect
        at the next commit, traversal options at the next query.  Results never depend on them."""
        _lib.check(self._L.qsmrt_scene_set_option(self._h, _lib.OPT[name], float(value)))

    def get_option(self, name: str) -> float:
        v = C.c_double()
        _lib.check(self._L.qsmrt_scene_get_option(self._h, _lib.OPT[name], C.byref(v)))
        return float(v.value)

    def counters(self) -> list:
        """The 16 fetch / lane counters of the last ``cast_rays`` launch made with ``set_option('counters', 1)``."""
        out = (C.c_uint64 * 16)()
        _lib.check(self._L.qsmrt_scene_get_counters(self._h, out))
        return list(out)

    def save(self, path: str, with_bvh: bool = True) -> None:
        """Write the scene (geometries; with ``with_bvh`` also the committed LBVH) to ``path`` -- the counterpart
        of the reference pickling its built search structures (``pyQSM/utils/io.py:44-60``)."""
        _lib.check(self._L.qsmrt_scene_save(self._h, str(path).encode(), _lib.SAVE_BVH if with_bvh else 0))

    @classmethod
    def load(cls, path: str, device=None, output_device=None) -> "RaycastingScene":
        """Scene from a file written by ``save``: committed at once if the file holds the BVH, otherwise rebuilt
        (deterministically: the identical tree) by the first query."""
        self = cls.__new__(cls)
        self._init_common(device, output_device)
        h = C.c_void_p()
        _lib.check(self._L.qsmrt_scene_load(self.device.index, str(path).encode(), C.byref(h)))
        self._h = h
        return self

    def stats(self) -> dict:
        st = _lib.Stats()
        _lib.check(self._L.qsmrt_get_stats(self._h, C.byref(st)))
        d = {k: getattr(st, k) for k, _ in st._fields_ if k not in ("scene_lo", "scene_hi", "reserved")}
        d["scene_lo"] = list(st.scene_lo)
        d["scene_hi"] = list(st.scene_hi)
        return d

    # -------------------------------------------------------------- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _rays(self, rays, name="rays"):
        y = _unwrap(rays)
        if y is None:
            y = np.asarray(rays, dtype=np.float32)
        if isinstance(y, np.ndarray):
            if y.dtype != np.float32:
                raise RuntimeError(f"{name} has dtype {y.dtype}, but it must be Float32")
            y = torch.from_numpy(np.ascontiguousarray(y))
        elif y.dtype != torch.float32:
            raise RuntimeError(f"{name} has dtype {y.dtype}, but it must be Float32")
        if y.ndim < 1 or y.shape[-1] != 6:
            raise RuntimeError(f"{name} has shape {tuple(y.shape)}, but the last dimension must be 6")
        return y

    def _out(self, t: torch.Tensor) -> torch.Tensor:
        if self.output_device.type == "cuda":
            return t
        return _to_host(t)

    # -------------------------------------------------------------- queries
    def cast_rays(self, rays, nthreads: int = 0, grid_width: int = 0, outputs=None) -> dict:
        """Closest hit per ray.  Keys: ``t_hit`` (inf on miss), ``geometry_ids``,
        ``primitive_ids`` (INVALID_ID on miss), ``primitive_uvs``, ``primitive_normals``.
        Rays shaped ``[H, W, 6]`` (or flat with ``grid_width=W``) are traversed in
        8x4 tiles; the results are the same either way.

        ``outputs``: ``None`` (default) -- all five keys; with CPU results, ``t_hit`` and
        ``primitive_ids`` are copied to the host at once and the other three on first
        access (``CastResult``).  ``"all"`` -- all five copied at once (a plain dict, as
        Open3D).  A tuple of key names -- only those are computed and returned."""
        r = self._rays(rays)
        shp = tuple(r.shape[:-1])
        n = int(np.prod(shp)) if shp else 1
        if outputs is None or outputs == "all":
            keys = CAST_KEYS
        else:
            keys = tuple(outputs)
            for k in keys:
                if k not in CAST_KEYS:
                    raise RuntimeError(f"cast_rays: unknown output {k!r}")
        host_out = self.output_device.type == "cpu"
        lazy = tuple(k for k in keys if k not in EAGER_KEYS) if (outputs is None and host_out and n >= (1 << 12)) else ()
        shape_of = lambda k: shp + _CAST_TAIL[k]
        if r.device.type == "cpu" and host_out:
            # host buffers in, host buffers out: chunked copy/compute overlap inside the C ABI
            r = r.contiguous()
            pin = n >= (1 << 16)
            eager = {k: torch.empty((n,) + _CAST_TAIL[k], dtype=_CAST_DTYPE[k], pin_memory=pin) for k in keys if k not in lazy}
            with torch.cuda.device(self.device):
                pending = {k: torch.empty((n,) + _CAST_TAIL[k], dtype=_CAST_DTYPE[k], device=self.device) for k in lazy}
            host_tab = (C.c_void_p * 5)(*[eager[k].data_ptr() if k in eager and n else None for k in CAST_KEYS])
            dev_tab = (C.c_void_p * 5)(*[pending[k].data_ptr() if k in pending and n else None for k in CAST_KEYS])
            _lib.check(self._L.qsmrt_cast_rays_host_split(self._h, _ptr(r), n, host_tab, dev_tab))
        else:
            with torch.cuda.device(self.device):
                r = r.to(self.device).contiguous()
                dev = {k: torch.empty((n,) + _CAST_TAIL[k], dtype=_CAST_DTYPE[k], device=self.device) for k in keys}
                ptrs = [_ptr(dev.get(k)) for k in CAST_KEYS]
                width = int(grid_width) if grid_width else (int(shp[-1]) if len(shp) >= 2 else 0)
                if width >= 8 and n % width == 0 and n // width >= 4:
                    # image / grid shaped batch (create_rays_pinhole output): 8x4 ray tiles per warp
                    _lib.check(self._L.qsmrt_cast_rays_2d(self._h, _ptr(r), width, n // width, *ptrs, self._stream()))
                else:
                    _lib.check(self._L.qsmrt_cast_rays(self._h, _ptr(r), n, *ptrs, self._stream()))
                eager = {k: self._out(dev[k]) for k in keys if k not in lazy}
                pending = {k: dev[k] for k in lazy}
        eager = {k: t.reshape(shape_of(k)) for k, t in eager.items()}
        if not lazy:
            return {k: eager[k] for k in keys}
        return CastResult({k: eager[k] for k in keys if k in eager}, {k: t.reshape(shape_of(k)) for k, t in pending.items()})

    def count_intersections(self, rays, nthreads: int = 0) -> torch.Tensor:
        """Number of intersections per ray (int32), Open3D dedup rule."""
        r = self._rays(rays)
        shp = tuple(r.shape[:-1])
        n = r.numel() // 6
        if r.device.type == "cpu" and self.output_device.type == "cpu" and n >= (1 << 16):
            r = r.contiguous()
            out = torch.empty(n, dtype=torch.int32, pin_memory=True)
            _lib.check(self._L.qsmrt_count_intersections_host(self._h, _ptr(r), n, _ptr(out)))
            return out.reshape(shp)
        with torch.cuda.device(self.device):
            r = r.to(self.device).contiguous()
            out = torch.empty(n, dtype=torch.int32, device=self.device)
            _lib.check(self._L.qsmrt_count_intersections(self._h, _ptr(r), n, _ptr(out), self._stream()))
            return self._out(out).reshape(shp)

    def test_occlusions(self, rays, tnear: float = 0.0, tfar: float = math.inf, nthreads: int = 0) -> torch.Tensor:
        """True where any triangle is hit with tnear < t <= tfar."""
        r = self._rays(rays)
        shp = tuple(r.shape[:-1])
        n = r.numel() // 6
        if r.device.type == "cpu" and self.output_device.type == "cpu" and n >= (1 << 16):
            r = r.contiguous()
            out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            _lib.check(self._L.qsmrt_test_occlusions_host(self._h, _ptr(r), n, float(tnear), float(tfar), _ptr(out)))
            return out.view(torch.bool).reshape(shp)
        with torch.cuda.device(self.device):
            r = r.to(self.device).contiguous()
            out = torch.empty(n, dtype=torch.uint8, device=self.device)
            _lib.check(self._L.qsmrt_test_occlusions(self._h, _ptr(r), n, float(tnear), float(tfar), _ptr(out), self._stream()))
            return self._out(out.to(torch.bool)).reshape(shp)

    def list_intersections(self, rays, nthreads: int = 0) -> dict:
        """All intersections per ray in CSR form (Open3D >= 0.18).  Keys:
        ``ray_splits`` [N+1], ``ray_ids``, ``t_hit``, ``geometry_ids``,
        ``primitive_ids``, ``primitive_uvs``; hits of a ray are sorted by t."""
        r = self._rays(rays)
        with torch.cuda.device(self.device):
            r = r.to(self.device).contiguous()
            n = r.numel() // 6
            splits = torch.empty(n + 1, dtype=torch.int64, device=self.device)
            total = C.c_int64()
            _lib.check(self._L.qsmrt_list_intersections_count(self._h, _ptr(r), n, _ptr(splits), C.byref(total), self._stream()))
            k = int(total.value)
            ray_ids = torch.empty(k, dtype=torch.int64, device=self.device)
            t_hit = torch.empty(k, dtype=torch.float32, device=self.device)
            gid = torch.empty(k, dtype=torch.uint32, device=self.device)
            pid = torch.empty(k, dtype=torch.uint32, device=self.device)
            uv = torch.empty(k, 2, dtype=torch.float32, device=self.device)
            _lib.check(self._L.qsmrt_list_intersections_fill(self._h, _ptr(r), n, _ptr(splits), _ptr(ray_ids), _ptr(t_hit),
                                                             _ptr(gid), _ptr(pid), _ptr(uv), self._stream()))
            return {"ray_splits": self._out(splits), "ray_ids": self._out(ray_ids), "t_hit": self._out(t_hit),
                    "geometry_ids": self._out(gid), "primitive_ids": self._out(pid), "primitive_uvs": self._out(uv)}

    def compute_occupancy(self, query_points, nthreads: int = 0, nsamples: int = 1) -> torch.Tensor:
        """1.0 inside / 0.0 outside (``ray_casting.py:69``): parity of the
        intersection count of a ray from each point.  ``nsamples == 1``: along
        (1,1,1), as Open3D's ``ComputeOccupancy``; odd ``nsamples > 1``: majority
        vote over that many fixed directions.

        KNOWN DIVERGENCE: Open3D draws the extra directions from ``std::mt19937(42)``;
        here they come from numpy ``default_rng(42)`` (the C++ stream is not
        reproduced).  For watertight meshes every direction gives the same parity, so
        the answers agree; for meshes with holes ``nsamples > 1`` can differ from
        Open3D's answer near the holes.  ``nsamples == 1`` (what the reference uses,
        ``ray_casting.py:69``) is not affected."""
        if nsamples < 1 or nsamples % 2 != 1:
            raise RuntimeError("compute_occupancy: nsamples must be odd and >= 1")
        p = _unwrap(query_points)
        if p is None:
            p = np.asarray(query_points, dtype=np.float32)
        if isinstance(p, np.ndarray):
            if p.dtype != np.float32:
                raise RuntimeError(f"query_points has dtype {p.dtype}, but it must be Float32")
            p = torch.from_numpy(np.ascontiguousarray(p))
        if p.ndim < 1 or p.shape[-1] != 3:
            raise RuntimeError(f"query_points has shape {tuple(p.shape)}, but the last dimension must be 3")
        with torch.cuda.device(self.device):
            p = p.to(self.device, torch.float32)
            saved, self.output_device = self.output_device, self.device
            try:
                votes = torch.zeros(p.shape[:-1], dtype=torch.int32, device=self.device)
                for d in _occupancy_directions(nsamples):
                    dirs = torch.tensor(d, dtype=torch.float32, device=self.device).expand_as(p)
                    cnt = self.count_intersections(torch.cat([p, dirs], dim=-1))
                    votes += (cnt % 2 == 1).to(torch.int32)
            finally:
                self.output_device = saved
            return self._out((2 * votes > nsamples).to(torch.float32))

    def _points(self, query_points):
        p = _unwrap(query_points)
        if p is None:
            p = np.asarray(query_points, dtype=np.float32)
        if isinstance(p, np.ndarray):
            if p.dtype != np.float32:
                raise RuntimeError(f"query_points has dtype {p.dtype}, but it must be Float32")
            p = torch.from_numpy(np.ascontiguousarray(p))
        elif p.dtype != torch.float32:
            raise RuntimeError(f"query_points has dtype {p.dtype}, but it must be Float32")
        if p.ndim < 1 or p.shape[-1] != 3:
            raise RuntimeError(f"query_points has shape {tuple(p.shape)}, but the last dimension must be 3")
        return p

    def compute_closest_points(self, query_points, nthreads: int = 0) -> dict:
        """Closest surface point per query point.  Keys as Open3D: ``points``,
        ``geometry_ids``, ``primitive_ids``, ``primitive_uvs``, ``primitive_normals``."""
        p = self._points(query_points)
        shp = tuple(p.shape[:-1])
        with torch.cuda.device(self.device):
            p = p.to(self.device).contiguous()
            n = p.numel() // 3
            pts = torch.empty(n, 3, dtype=torch.float32, device=self.device)
            gid = torch.empty(n, dtype=torch.uint32, device=self.device)
            pid = torch.empty(n, dtype=torch.uint32, device=self.device)
            uv = torch.empty(n, 2, dtype=torch.float32, device=self.device)
            nrm = torch.empty(n, 3, dtype=torch.float32, device=self.device)
            _lib.check(self._L.qsmrt_closest_points(self._h, _ptr(p), n, _ptr(pts), None, _ptr(gid), _ptr(pid), _ptr(uv), _ptr(nrm),
                                                    self._stream()))
            return {"points": self._out(pts).reshape(shp + (3,)), "geometry_ids": self._out(gid).reshape(shp),
                    "primitive_ids": self._out(pid).reshape(shp), "primitive_uvs": self._out(uv).reshape(shp + (2,)),
                    "primitive_normals": self._out(nrm).reshape(shp + (3,))}

    def compute_distance(self, query_points, nthreads: int = 0) -> torch.Tensor:
        """Unsigned distance to the closest surface point."""
        p = self._points(query_points)
        shp = tuple(p.shape[:-1])
        with torch.cuda.device(self.device):
            p = p.to(self.device).contiguous()
            n = p.numel() // 3
            d = torch.empty(n, dtype=torch.float32, device=self.device)
            _lib.check(self._L.qsmrt_closest_points(self._h, _ptr(p), n, None, _ptr(d), None, None, None, None, self._stream()))
            return self._out(d).reshape(shp)

    def compute_signed_distance(self, query_points, nthreads: int = 0, nsamples: int = 1) -> torch.Tensor:
        """Distance, negative inside (``ray_casting.py:250,255``): the sign is
        ``compute_occupancy`` (odd intersection count along (1,1,1); majority of
        ``nsamples`` directions when ``nsamples > 1``)."""
        if nsamples < 1 or nsamples % 2 != 1:
            raise RuntimeError("compute_signed_distance: nsamples must be odd and >= 1")
        p = self._points(query_points)
        shp = tuple(p.shape[:-1])
        with torch.cuda.device(self.device):
            p = p.to(self.device).contiguous()
            n = p.numel() // 3
            d = torch.empty(n, dtype=torch.float32, device=self.device)
            if nsamples == 1:
                _lib.check(self._L.qsmrt_signed_distance(self._h, _ptr(p), n, _ptr(d), self._stream()))
            else:
                saved, self.output_device = self.output_device, self.device
                try:
                    d = self.compute_distance(p.reshape(-1, 3)).reshape(-1)
                    inside = self.compute_occupancy(p.reshape(-1, 3), nsamples=nsamples).reshape(-1) > 0
                finally:
                    self.output_device = saved
                d = torch.where(inside, -d, d)
            return self._out(d).reshape(shp)

    def mark_hit_primitives(self, ans: dict, vertices: bool = False):
        """Device-side form of ``ray_casting.py:285-292``: uint8 flags of the
        triangles (scene order) that own a closest hit -- ``np.unique(prim_ids)`` as a
        mask -- and, with ``vertices=True``, also of their corner vertices
        (``hit_vert_ids = np.unique(triangles[prim_ids])``): returns ``tri`` or
        ``(tri, vert)``."""
        with torch.cuda.device(self.device):
            gid = ans["geometry_ids"].to(self.device).contiguous().reshape(-1)
            pid = ans["primitive_ids"].to(self.device).contiguous().reshape(-1)
            self.commit()
            nt = int(self.stats()["num_triangles"])
            nv = sum(self.geometry_size(g)[0] for g in range(int(self.stats()["num_geometries"]))) if vertices else 0
            tri = torch.zeros(max(nt, 1), dtype=torch.uint8, device=self.device)
            vert = torch.zeros(max(nv, 1), dtype=torch.uint8, device=self.device) if vertices else None
            _lib.check(self._L.qsmrt_mark_hit_primitives(self._h, _ptr(gid), _ptr(pid), pid.numel(), _ptr(tri), _ptr(vert), self._stream()))
            if vertices:
                return self._out(tri[:nt]), self._out(vert[:nv])
            return self._out(tri[:nt])

    def geometry_size(self, geometry_id: int):
        """(number of vertices, number of triangles) of a registered geometry."""
        nv, nt = C.c_uint64(), C.c_uint64()
        _lib.check(self._L.qsmrt_geometry_size(self._h, int(geometry_id), C.byref(nv), C.byref(nt)))
        return int(nv.value), int(nt.value)

    def vertex_exposure(self, tri_counts: torch.Tensor) -> torch.Tensor:
        """Per-vertex exposure from per-triangle exposure (scene order): every vertex
        receives the counts of the triangles it is a corner of -- BASELINE config 2's
        "sunlight exposure per leaf vertex", the count form of ``ray_casting.py:289-292``."""
        with torch.cuda.device(self.device):
            self.commit()
            st = self.stats()
            nv = sum(self.geometry_size(g)[0] for g in range(int(st["num_geometries"])))
            tc = tri_counts.to(self.device).contiguous().view(torch.int32)
            if tc.numel() != int(st["num_triangles"]):
                raise RuntimeError(f"tri_counts has {tc.numel()} entries, the scene {int(st['num_triangles'])} triangles")
            out = torch.zeros(max(nv, 1), dtype=torch.int32, device=self.device)
            _lib.check(self._L.qsmrt_vertex_exposure(self._h, _ptr(tc), _ptr(out), self._stream()))
            return out[:nv]

    # ------------------------------------------------------- ray generators
    @staticmethod
    def create_rays_pinhole(*args, **kwargs) -> torch.Tensor:
        """``create_rays_pinhole(intrinsic_matrix, extrinsic_matrix, width_px, height_px)`` or
        ``create_rays_pinhole(fov_deg, center, eye, up, width_px, height_px)``
        -> float32 ``[height_px, width_px, 6]`` (origins = eye, directions not
        normalised), as Open3D's ``CreateRaysPinhole``.  ``device=`` keeps the
        rays on the GPU; the default returns a CPU tensor like Open3D."""
        device = kwargs.pop("device", None)
        names_a = ("intrinsic_matrix", "extrinsic_matrix", "width_px", "height_px")
        names_b = ("fov_deg", "center", "eye", "up", "width_px", "height_px")
        if "fov_deg" in kwargs or len(args) == 6 or (len(args) > 0 and np.ndim(_np64(args[0])) == 0):
            a = dict(zip(names_b, args)); a.update(kwargs)
            w, h = int(a["width_px"]), int(a["height_px"])
            K, E = _fov_to_matrices(float(a["fov_deg"]), _np64(a["center"]), _np64(a["eye"]), _np64(a["up"]), w, h)
        else:
            a = dict(zip(names_a, args)); a.update(kwargs)
            w, h = int(a["width_px"]), int(a["height_px"])
            K, E = _np64(a["intrinsic_matrix"]), _np64(a["extrinsic_matrix"])
            if K.shape != (3, 3):
                raise RuntimeError(f"intrinsic_matrix has shape {K.shape}, but it must be (3, 3)")
            if E.shape != (4, 4):
                raise RuntimeError(f"extrinsic_matrix has shape {E.shape}, but it must be (4, 4)")
        L = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("create_rays_pinhole needs a CUDA device; there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None or torch.device(device).type != "cuda" \
            else torch.device(device)
        with torch.cuda.device(dev):
            rays = torch.empty(h, w, 6, dtype=torch.float32, device=dev)
            Kc = (C.c_double * 9)(*np.ascontiguousarray(K).reshape(-1))
            Ec = (C.c_double * 16)(*np.ascontiguousarray(E).reshape(-1))
            _lib.check(L.qsmrt_gen_pinhole_rays(_ptr(rays), w, h, Kc, Ec, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        if device is not None and torch.device(device).type == "cuda":
            return rays
        return _to_host(rays)


def _np64(x):
    y = _unwrap(x)
    if isinstance(y, torch.Tensor):
        y = y.detach().cpu().numpy()
    return np.asarray(x if y is None else y, dtype=np.float64)


def _fov_to_matrices(fov_deg, center, eye, up, w, h):
    """Open3D CreateRaysPinhole(fov_deg, center, eye, up, w, h): focal length
    0.5*w/tan(0.5*fov); camera rows R2 = normalize(center-eye),
    R0 = normalize(up x R2), R1 = R2 x R0; t = -R eye."""
    f = 0.5 * w / math.tan(0.5 * math.radians(fov_deg))
    K = np.array([[f, 0, 0.5 * w], [0, f, 0.5 * h], [0, 0, 1]], dtype=np.float64)
    R = np.zeros((3, 3))
    R[1] = up / np.linalg.norm(up)
    R[2] = center - eye
    R[2] /= np.linalg.norm(R[2])
    R[0] = np.cross(R[1], R[2])
    R[0] /= np.linalg.norm(R[0])
    R[1] = np.cross(R[2], R[0])
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = -R @ eye
    return K, E
