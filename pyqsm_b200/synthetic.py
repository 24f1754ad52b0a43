"""Deterministic synthetic meshes and ray sets for the BASELINE.json configs
(SURVEY.md section 8d).  Host-side numpy: these only *shape* the inputs the
reference would hand to ``RaycastingScene`` -- cylinder-QSM meshes
(``pyQSM/geometry/point_cloud_processing.py:274-279`` create_cylinder) and
reconstructed canopy meshes -- and the parallel-ray grids of
``pyQSM/viz/ray_casting.py:159-165``.  No ray/triangle arithmetic lives here.
"""
from __future__ import annotations

import math

import numpy as np


# ----------------------------------------------------------------- meshes
def cylinder_mesh(radius=1.0, height=2.0, resolution=20, split=4):
    """Open3D ``TriangleMesh.create_cylinder`` topology: axis z, centred at the
    origin; 2 cap centres + (split+1) rings of ``resolution`` vertices;
    2*resolution cap triangles + 2*resolution*split side triangles
    (102 vertices / 200 triangles at the defaults)."""
    v = np.zeros((resolution * (split + 1) + 2, 3), np.float64)
    v[0] = (0, 0, height * 0.5)
    v[1] = (0, 0, -height * 0.5)
    step = 2.0 * math.pi / resolution
    hstep = height / split
    for i in range(split + 1):
        for j in range(resolution):
            th = step * j
            v[2 + resolution * i + j] = (math.cos(th) * radius, math.sin(th) * radius, height * 0.5 - hstep * i)
    t = []
    for j in range(resolution):
        j1 = (j + 1) % resolution
        t.append((0, 2 + j, 2 + j1))
        b = 2 + resolution * split
        t.append((1, b + j1, b + j))
    for i in range(split):
        b1 = 2 + resolution * i
        b2 = b1 + resolution
        for j in range(resolution):
            j1 = (j + 1) % resolution
            t.append((b2 + j, b1 + j1, b1 + j))
            t.append((b2 + j, b2 + j1, b1 + j1))
    return v.astype(np.float32), np.asarray(t, np.uint32)


def box_mesh(lo=(0, 0, 0), hi=(1, 1, 1)):
    """Axis-aligned box, 8 vertices / 12 triangles, faces split on a diagonal."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    v = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    i = lambda x, y, z: x * 4 + y * 2 + z
    q = lambda a, b, c, d: [(a, b, c), (a, c, d)]
    t = (q(i(0, 0, 0), i(1, 0, 0), i(1, 1, 0), i(0, 1, 0)) + q(i(0, 0, 1), i(1, 0, 1), i(1, 1, 1), i(0, 1, 1)) +
         q(i(0, 0, 0), i(1, 0, 0), i(1, 0, 1), i(0, 0, 1)) + q(i(0, 1, 0), i(1, 1, 0), i(1, 1, 1), i(0, 1, 1)) +
         q(i(0, 0, 0), i(0, 1, 0), i(0, 1, 1), i(0, 0, 1)) + q(i(1, 0, 0), i(1, 1, 0), i(1, 1, 1), i(1, 0, 1)))
    return v.astype(np.float32), np.asarray(t, np.uint32)


def _rotation_to(axis):
    """Rotation taking +z onto ``axis`` (Rodrigues)."""
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    z = np.array([0.0, 0.0, 1.0])
    v = np.cross(z, a)
    s, c = np.linalg.norm(v), float(z @ a)
    if s < 1e-12:
        return np.eye(3) if c > 0 else np.diag([1.0, -1.0, -1.0])
    k = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]]) / s
    return np.eye(3) + s * k + (1 - c) * (k @ k)


def qsm_tree_mesh(seed=1, n_cylinders=250, resolution=20, split=4):
    """C1: cylinder-QSM tree.  Trunk r=0.25 m h=6 m; 3 children per segment,
    radius x0.7, length x0.75, branch angle U(25,60) deg, azimuth U(0,2pi),
    breadth-first to ``n_cylinders`` (250 -> 25 500 vertices / 50 000 triangles)."""
    rng = np.random.default_rng(seed)
    cv, ct = cylinder_mesh(1.0, 1.0, resolution, split)
    segs = [(np.zeros(3), np.array([0.0, 0.0, 1.0]), 0.25, 6.0)]     # base, axis, radius, length
    verts, tris = [], []
    q = 0
    while q < len(segs) and q < n_cylinders:
        base, axis, r, L = segs[q]
        q += 1
        R = _rotation_to(axis)
        v = cv.astype(np.float64) * np.array([r, r, L])
        v[:, 2] += 0.5 * L
        v = v @ R.T + base
        tris.append(ct + np.uint32(len(verts) * cv.shape[0]))
        verts.append(v)
        tip = base + axis * L
        for _ in range(3):
            ang = math.radians(rng.uniform(25.0, 60.0))
            az = rng.uniform(0.0, 2.0 * math.pi)
            local = np.array([math.sin(ang) * math.cos(az), math.sin(ang) * math.sin(az), math.cos(ang)])
            segs.append((tip, R @ local, r * 0.7, L * 0.75))
    return np.concatenate(verts).astype(np.float32), np.concatenate(tris).astype(np.uint32)


def canopy_mesh(seed=2, n_leaves=1_000_000, semi_axes=(6.0, 6.0, 4.0), center=(0.0, 0.0, 10.0),
                shell=(0.6, 1.0), leaf_size=(0.05, 0.03)):
    """C2/C5: leaf-soup canopy.  ``n_leaves`` quads (2 triangles, 4 unshared
    vertices each; 1M -> 2M triangles) with centres uniform in an ellipsoidal
    shell of 12 x 12 x 8 m, 5 x 3 cm leaves, normals uniform on the sphere
    (spherical leaf-angle distribution), random in-plane rotation."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_leaves, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    u = rng.uniform(shell[0] ** 3, shell[1] ** 3, size=(n_leaves, 1)) ** (1.0 / 3.0)
    c = d * u * np.asarray(semi_axes) + np.asarray(center)
    nrm = rng.normal(size=(n_leaves, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    h = rng.normal(size=(n_leaves, 3))
    a = np.cross(nrm, h)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = np.cross(nrm, a)
    a *= 0.5 * leaf_size[0]
    b *= 0.5 * leaf_size[1]
    v = np.stack([c - a - b, c + a - b, c + a + b, c - a + b], axis=1).reshape(-1, 3)
    base = (np.arange(n_leaves, dtype=np.uint32) * 4)[:, None]
    t = np.concatenate([base + np.array([0, 1, 2], np.uint32), base + np.array([0, 2, 3], np.uint32)], axis=1).reshape(-1, 3)
    return v.astype(np.float32), t.astype(np.uint32)


def plot_mesh(seed=3, n_canopies=5, leaves_per_canopy=1_000_000, pitch=14.0):
    """C3/C4: a plot of canopies tiled on a square grid (5 -> 10M triangles;
    25 at pitch 40 m -> 50M)."""
    side = int(math.ceil(math.sqrt(n_canopies)))
    vs, ts, off = [], [], 0
    for k in range(n_canopies):
        v, t = canopy_mesh(seed * 1000 + k, leaves_per_canopy)
        v = v + np.array([(k % side) * pitch, (k // side) * pitch, 0.0], np.float32)
        vs.append(v)
        ts.append(t + np.uint32(off))
        off += v.shape[0]
    return np.concatenate(vs), np.concatenate(ts)


# -------------------------------------------------------------------- rays
def sun_direction(elevation_deg, azimuth_deg):
    """Unit vector of travel of sunlight (pointing down, away from the sun)."""
    el, az = math.radians(elevation_deg), math.radians(azimuth_deg)
    return -np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])


def parallel_ray_grid(lo, hi, direction, nu, nv, margin=0.05):
    """Grid of parallel rays covering the scene AABB [lo, hi] as seen along
    ``direction``: returns float32 (origin0, du, dv, dir) so that ray (i, j) is
    origin0 + i*du + j*dv.  Origins sit on a plane perpendicular to the
    direction, outside the bounding sphere."""
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    d = np.asarray(direction, np.float64)
    d = d / np.linalg.norm(d)
    ref = np.array([0.0, 0.0, 1.0]) if abs(d[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    u = np.cross(ref, d)
    u /= np.linalg.norm(u)
    v = np.cross(d, u)
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    ctr = 0.5 * (lo + hi)
    pu, pv = (corners - ctr) @ u, (corners - ctr) @ v
    su, sv = (pu.max() - pu.min()) * (1 + margin), (pv.max() - pv.min()) * (1 + margin)
    cu, cv = 0.5 * (pu.max() + pu.min()), 0.5 * (pv.max() + pv.min())
    radius = 0.5 * np.linalg.norm(hi - lo) * 1.1
    du, dv = u * (su / nu), v * (sv / nv)
    o0 = ctr + u * (cu - 0.5 * su) + v * (cv - 0.5 * sv) + 0.5 * du + 0.5 * dv - d * radius
    f = lambda a: np.asarray(a, np.float32)
    return f(o0), f(du), f(dv), f(d)


def materialize_grid(o0, du, dv, d, nu, nv):
    """numpy float32 rays [nv*nu, 6] of a parallel grid (i fastest).  The CUDA
    generator fuses the multiply-adds, so values can differ by an ulp: parity
    tests feed both sides the device-generated rays."""
    i = np.arange(nu, dtype=np.float32)[None, :, None]
    j = np.arange(nv, dtype=np.float32)[:, None, None]
    o = (o0[None, None, :] + i * du[None, None, :] + j * dv[None, None, :]).astype(np.float32)
    r = np.empty((nv, nu, 6), np.float32)
    r[..., :3] = o
    r[..., 3:] = d
    return r.reshape(-1, 6)


def hemisphere_sweep(n_elevation=8, n_azimuth=8):
    """C2: 8 elevations {10..80 deg} x 8 azimuths {0..315 deg}."""
    els = np.linspace(10.0, 80.0, n_elevation)
    azs = np.arange(n_azimuth) * (360.0 / n_azimuth)
    return [(float(e), float(a)) for e in els for a in azs]


def random_rays(lo, hi, n, seed=0):
    """Incoherent rays: origins uniform in an inflated AABB, directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    ext = hi - lo
    o = rng.uniform(lo - 0.25 * ext, hi + 0.25 * ext, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1).astype(np.float32)
