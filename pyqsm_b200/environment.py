"""Environmental ray-casting drivers: sunlight angle, rain angle and sky
(cloud-cover / gap-fraction) simulations over a ``RaycastingScene``.

The reference only states the intent -- ``README.md:127`` ("simulate ...
sunlight angle, cloud cover and rain angle"), ``data/notes/methods.md:16,53-55``
and ``data/notes/epiphyte_isolation_methods.md:17,45`` ("Ray casting: Parallel
rays from nadir") -- and its only parallel-ray code is the 10 x 10 vertical
grid of ``pyQSM/viz/ray_casting.py:159-165``.  These drivers are that pattern
at scale: rays are generated inside the
traversal kernel and the per-triangle / per-point results are reduced on the
device, so neither the 24 B/ray of input nor the 32 B/ray of ``cast_rays``
output exists.  Each function equals a composition of the ``RaycastingScene``
queries (the tests check exactly that).

Multi-GPU: pass ``shard=(rank, world)`` to take this rank's share of the
angles / directions and all-reduce the integer results over
``torch.distributed`` (no collective during traversal).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib, synthetic as syn
from .distributed import allreduce_sum, shard_angles, shard_range


def _f3(x):
    return (C.c_float * 3)(*[float(v) for v in x])


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def sun_exposure(scene, angles, grid=(4000, 4000), margin=0.05, shard=None, per_angle=False, per_vertex=False):
    """Sunlit-ray counts per triangle for a sweep of solar ``angles`` =
    [(elevation_deg, azimuth_deg), ...].

    For every angle a ``grid = (nu, nv)`` of parallel rays covering the scene
    bounds (as seen from the sun) is cast; a triangle's count is the number of
    rays whose *closest* hit it owns, i.e. its sunlit projected area in units
    of one ray cell.  The whole sweep is ONE kernel launch
    (``qsmrt_sun_exposure_sweep``): rays are generated in the kernel and hits
    are added per triangle with warp-aggregated atomics.  Returns ``{"counts":
    int32 [T] (or [A, T] with per_angle), "cell_area": [A] m^2 per ray,
    "rays": total rays cast}`` on the scene's device; with ``per_vertex`` also
    ``"vertex_counts"`` int32 [V] (or [A, V]): every vertex receives the counts
    of the triangles it is a corner of (BASELINE config 2, "sunlight exposure
    per leaf vertex"; ``ray_casting.py:289-292``).  Triangles and vertices are
    in scene order (geometries in the order added)."""
    L = _lib.load()
    scene.commit()
    st = scene.stats()
    ntri = int(st["num_triangles"])
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    angles = list(angles)
    mine = shard_angles(list(enumerate(angles)), *shard) if shard else list(enumerate(angles))
    nu, nv = int(grid[0]), int(grid[1])
    dev = scene.device
    with torch.cuda.device(dev):
        rows = len(angles) if per_angle else 1
        counts = torch.zeros(rows, max(ntri, 1), dtype=torch.int32, device=dev)
        cell = torch.zeros(len(angles), dtype=torch.float64)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        grids = np.empty((len(mine), 12), np.float32)
        for j, (k, (el, az)) in enumerate(mine):
            o0, du, dv, d = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), nu, nv, margin)
            cell[k] = float(np.linalg.norm(du.astype(np.float64)) * np.linalg.norm(dv.astype(np.float64)))
            grids[j] = np.concatenate([o0, du, dv, d])
        if per_angle:
            # one launch per contiguous run of this rank's angles (row a of the launch lands in row k0 + a)
            j = 0
            while j < len(mine):
                e = j
                while e + 1 < len(mine) and mine[e + 1][0] == mine[e][0] + 1:
                    e += 1
                blk = np.ascontiguousarray(grids[j:e + 1])
                _lib.check(L.qsmrt_sun_exposure_sweep(scene._h, e + 1 - j, blk.ctypes.data_as(C.POINTER(C.c_float)), nu, nv,
                                                      _ptr(counts[mine[j][0]]), counts.shape[1], stream))
                j = e + 1
        elif len(mine):
            _lib.check(L.qsmrt_sun_exposure_sweep(scene._h, len(mine), grids.ctypes.data_as(C.POINTER(C.c_float)), nu, nv,
                                                  _ptr(counts[0]), 0, stream))
        if shard:
            allreduce_sum(counts)
            cell_dev = cell.to(dev)
            allreduce_sum(cell_dev)
            cell = cell_dev.cpu()
        counts = counts[:, :ntri]
        out = {"counts": counts if per_angle else counts[0], "cell_area": cell, "rays": nu * nv * len(angles)}
        if per_vertex:
            vc = torch.stack([scene.vertex_exposure(counts[a]) for a in range(counts.shape[0])])
            out["vertex_counts"] = vc if per_angle else vc[0]
    return out


def rain_interception(scene, angle_from_vertical_deg=20.0, azimuth_deg=0.0, grid=(10000, 10000), margin=0.05,
                      chunk_rows=2000, shard=None):
    """Rain at a slant: a ``grid`` of parallel rays ``angle_from_vertical_deg``
    off nadir.  Returns ``{"intersections": int64 histogram of the per-ray
    intersection counts, "intercepted_fraction": share of rays that hit
    anything, "mean_layers": mean number of surfaces crossed}`` using
    ``count_intersections`` (how many canopy layers a drop would cross)."""
    L = _lib.load()
    scene.commit()
    st = scene.stats()
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    nu, nv = int(grid[0]), int(grid[1])
    o0, du, dv, d = syn.parallel_ray_grid(lo, hi, syn.sun_direction(90.0 - angle_from_vertical_deg, azimuth_deg), nu, nv, margin)
    rows = shard_range(nv, *shard) if shard else (0, nv)
    dev = scene.device
    hist = torch.zeros(256, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        buf = torch.empty(min(chunk_rows, nv) * nu, 6, dtype=torch.float32, device=dev)
        cnt = torch.empty(min(chunk_rows, nv) * nu, dtype=torch.int32, device=dev)
        for r0 in range(rows[0], rows[1], chunk_rows):
            nr = min(chunk_rows, rows[1] - r0)
            o = (o0.astype(np.float64) + r0 * dv.astype(np.float64)).astype(np.float32)
            _lib.check(L.qsmrt_gen_parallel_rays(_ptr(buf), nu, nr, _f3(o), _f3(du), _f3(dv), _f3(d), stream))
            _lib.check(L.qsmrt_count_intersections(scene._h, _ptr(buf), nu * nr, _ptr(cnt), stream))
            hist += torch.bincount(cnt[: nu * nr].clamp(max=255).to(torch.int64), minlength=256)
        if shard:
            allreduce_sum(hist)
    total = int(hist.sum())
    layers = torch.arange(256, device=dev, dtype=torch.float64)
    return {"intersections": hist, "intercepted_fraction": float(1.0 - hist[0].item() / max(total, 1)),
            "mean_layers": float((hist.to(torch.float64) * layers).sum().item() / max(total, 1)), "rays": total}


def sky_gap_fraction(scene, points, normals=None, n_dirs=1000, seed=5, offset=1e-4, shard=None, point_base=0):
    """Diffuse-sky (cloud cover) visibility: for each query point the share of
    ``n_dirs`` directions, uniform over the upper hemisphere about +z, along
    which no triangle is hit (any-hit occlusion from ``point + offset *
    normal``).  Directions come from a counter-based hash of (seed, point,
    k), so a point sees the same sample however the work is sharded;
    ``point_base`` is the index of ``points[0]`` in that sample (a block cut
    out of a larger point set keeps its directions).
    Returns float32 ``[n_points]`` on the scene's device."""
    L = _lib.load()
    dev = scene.device
    scene.commit()
    with torch.cuda.device(dev):
        p = torch.as_tensor(points, dtype=torch.float32).to(dev).contiguous().reshape(-1, 3)
        nrm = None if normals is None else torch.as_tensor(normals, dtype=torch.float32).to(dev).contiguous().reshape(-1, 3)
        free = torch.zeros(max(p.shape[0], 1), dtype=torch.int32, device=dev)
        b, e = shard_range(int(n_dirs), *shard) if shard else (0, int(n_dirs))
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if e > b and p.shape[0]:
            _lib.check(L.qsmrt_sky_visibility(scene._h, _ptr(p), _ptr(nrm), p.shape[0], int(point_base), int(seed), float(offset), b, e - b,
                                              _ptr(free), stream))
        if shard:
            allreduce_sum(free)
        return free[: p.shape[0]].to(torch.float32) / float(n_dirs)


def peel_projection(scene, direction=(0.0, 0.0, -1.0), grid=(2000, 2000), margin=0.05, max_layers=256):
    """"Raycasting projection" of the reference's methods notes
    (``data/notes/methods.md:53-55``: "progressively casting rays,
    removing/summing the area of interception regions and repeating until all
    mesh components have been removed"; the ``surf_2d`` branch of
    ``ray_casting.py:285-301`` iterated).  Parallel rays along ``direction``
    (default: from nadir); each layer sums the area of the triangles that own
    a closest hit and removes them.  Returns ``{"layers": [(n_triangles,
    area_3d, area_projected), ...], "area_3d", "area_projected", "layer_of":
    int32 [T] (-1 = never seen)}`` -- overlapping surfaces are counted
    separately, unlike a single projection."""
    L = _lib.load()
    scene.commit()
    st = scene.stats()
    ntri = int(st["num_triangles"])
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    nu, nv = int(grid[0]), int(grid[1])
    o0, du, dv, d = syn.parallel_ray_grid(lo, hi, np.asarray(direction, np.float64), nu, nv, margin)
    dev = scene.device
    with torch.cuda.device(dev):
        layer_of = torch.full((max(ntri, 1),), -1, dtype=torch.int32, device=dev)
        stats = (C.c_double * (3 * max_layers))()
        nl = C.c_int(0)
        _lib.check(L.qsmrt_peel_projection(scene._h, nu, nv, _f3(o0), _f3(du), _f3(dv), _f3(d), int(max_layers), _ptr(layer_of),
                                           stats, C.byref(nl), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    layers = [(int(stats[3 * k]), float(stats[3 * k + 1]), float(stats[3 * k + 2])) for k in range(nl.value)]
    return {"layers": layers, "area_3d": sum(x[1] for x in layers), "area_projected": sum(x[2] for x in layers),
            "layer_of": layer_of[:ntri]}


def hemisphere_rays(points, normals=None, n_dirs=16, seed=5, offset=1e-4, dir_begin=0, device=None, point_base=0):
    """The exact rays ``sky_gap_fraction`` traces, materialised: float32 ``[n_points * n_dirs, 6]``."""
    L = _lib.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(dev):
        p = torch.as_tensor(points, dtype=torch.float32).to(dev).contiguous().reshape(-1, 3)
        nrm = None if normals is None else torch.as_tensor(normals, dtype=torch.float32).to(dev).contiguous().reshape(-1, 3)
        rays = torch.empty(p.shape[0] * int(n_dirs), 6, dtype=torch.float32, device=dev)
        if rays.numel():
            _lib.check(L.qsmrt_gen_hemisphere_rays(_ptr(rays), _ptr(p), _ptr(nrm), p.shape[0], int(point_base), int(seed), float(offset),
                                                   int(dir_begin), int(n_dirs), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return rays
