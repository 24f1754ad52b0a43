"""GPU tests of the environmental drivers: each fused
driver must equal the composition of RaycastingScene queries it stands for,
and that composition is checked against the oracle on the same rays."""
import ctypes as C

import numpy as np
import pytest
import torch

from pyqsm_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tree(oracle_mod):
    from pyqsm_b200 import RaycastingScene
    v, t = syn.qsm_tree_mesh(seed=3, n_cylinders=80)
    v2, t2 = syn.box_mesh((-3, -3, 0), (-2, -2, 1))                 # a second geometry: exposure uses scene order
    g = RaycastingScene(output_device="cuda")
    o = oracle_mod.OracleScene()
    for s in (g, o):
        s.add_triangles(v, t)
        s.add_triangles(v2, t2)
    return g, o, t.shape[0] + t2.shape[0], t.shape[0]


def test_sun_exposure_equals_cast_rays(tree):
    from pyqsm_b200 import environment as env, _lib
    g, o, ntri, n0 = tree
    L = _lib.load()
    angles = [(30.0, 40.0), (75.0, 200.0), (10.0, 315.0)]
    nu, nv = 257, 190
    res = env.sun_exposure(g, angles, grid=(nu, nv), per_angle=True)
    assert res["counts"].shape == (3, ntri) and res["counts"].dtype == torch.int32
    st = g.stats()
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
    for k, (el, az) in enumerate(angles):
        grid = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), nu, nv)
        rays = torch.empty(nu * nv, 6, dtype=torch.float32, device="cuda")
        _lib.check(L.qsmrt_gen_parallel_rays(C.c_void_p(rays.data_ptr()), nu, nv, F3(grid[0]), F3(grid[1]), F3(grid[2]), F3(grid[3]), None))
        torch.cuda.synchronize()
        ref = o.cast_rays(rays.cpu().numpy(), 1)                          # oracle on the device-generated rays
        hit = ref["primitive_ids"] != 0xFFFFFFFF
        tri = ref["primitive_ids"][hit].astype(np.int64) + np.where(ref["geometry_ids"][hit] == 1, n0, 0)
        expect = np.bincount(tri, minlength=ntri)
        assert hit.sum() > 500
        assert np.array_equal(res["counts"][k].cpu().numpy(), expect), f"angle {k}"
        assert abs(res["cell_area"][k].item() - np.linalg.norm(grid[1]) * np.linalg.norm(grid[2])) < 1e-9
    summed = env.sun_exposure(g, angles, grid=(nu, nv))
    assert torch.equal(summed["counts"], res["counts"].sum(0).to(torch.int32))


def test_sky_gap_fraction_equals_occlusion(tree):
    from pyqsm_b200 import environment as env
    g, o, ntri, n0 = tree
    rng = np.random.default_rng(4)
    pts = rng.uniform([-4, -4, 0.1], [6, 6, 12], size=(300, 3)).astype(np.float32)
    nrm = rng.normal(size=(300, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    n_dirs = 64
    gap = env.sky_gap_fraction(g, pts, nrm, n_dirs=n_dirs, seed=11, offset=1e-3)
    rays = env.hemisphere_rays(pts, nrm, n_dirs=n_dirs, seed=11, offset=1e-3)
    assert rays.shape == (300 * n_dirs, 6)
    r = rays.cpu().numpy()
    d = r[:, 3:]
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, rtol=2e-6)
    assert np.all(d[:, 2] >= 0) and 0.4 < d[:, 2].mean() < 0.6           # uniform in solid angle: E[z] = 1/2
    np.testing.assert_allclose(r[:, :3].reshape(300, n_dirs, 3)[:, 0], pts + np.float32(1e-3) * nrm, rtol=0, atol=1e-6)
    occ_g = g.test_occlusions(rays).cpu().numpy()
    occ_o = o.test_occlusions(r, mode=1)
    assert np.array_equal(occ_g, occ_o)
    expect = 1.0 - occ_o.reshape(300, n_dirs).mean(1)
    np.testing.assert_allclose(gap.cpu().numpy(), expect.astype(np.float32), rtol=0, atol=1e-7)
    assert 0.05 < (1 - expect).mean() < 0.95
    # direction subsets are disjoint parts of the same sample: shard sums equal the whole
    from pyqsm_b200 import _lib
    L = _lib.load()
    p_d, n_d = torch.from_numpy(pts).cuda(), torch.from_numpy(nrm).cuda()
    parts = torch.zeros(300, dtype=torch.int32, device="cuda")
    for b, c in ((0, 20), (20, 30), (50, 14)):
        _lib.check(L.qsmrt_sky_visibility(g._h, C.c_void_p(p_d.data_ptr()), C.c_void_p(n_d.data_ptr()), 300, 0, 11, 1e-3, b, c,
                                          C.c_void_p(parts.data_ptr()), None))
    torch.cuda.synchronize()
    assert torch.equal(parts.to(torch.float32) / n_dirs, gap)


def test_sky_point_order_does_not_change_the_answer(tree):
    """qsmrt_sky_visibility works large point sets in Morton order (a permutation the kernel reads); the sample of a
    point and where its count lands do not depend on it."""
    from pyqsm_b200 import environment as env
    g, o, ntri, n0 = tree
    rng = np.random.default_rng(9)
    pts = rng.uniform([-4, -4, 0.1], [6, 6, 12], size=(6000, 3)).astype(np.float32)      # >= 4096 points: the sorted path
    pts[17] = np.nan                                                                      # a NaN point sorts somewhere and sees nothing
    try:
        g.set_option("point_order", 0)
        a = env.sky_gap_fraction(g, pts, None, n_dirs=32, seed=3)
        g.set_option("point_order", 1)
        b = env.sky_gap_fraction(g, pts, None, n_dirs=32, seed=3)
    finally:
        g.set_option("point_order", 1)
    assert torch.equal(a, b) and 0.05 < float(b[torch.isfinite(b)].mean()) < 0.999
    blk = slice(3000, 3040)                                                               # a block, materialised with its place in the sample
    rays = env.hemisphere_rays(pts[blk], None, n_dirs=32, seed=3, point_base=3000)
    occ = o.test_occlusions(rays.cpu().numpy(), mode=1).reshape(40, 32)
    np.testing.assert_allclose(b[blk].cpu().numpy(), (1.0 - occ.mean(1)).astype(np.float32), rtol=0, atol=1e-7)


def test_rain_interception_equals_count(tree):
    from pyqsm_b200 import environment as env
    g, o, ntri, n0 = tree
    res = env.rain_interception(g, angle_from_vertical_deg=20.0, azimuth_deg=30.0, grid=(300, 240), chunk_rows=100)
    st = g.stats()
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    grid = syn.parallel_ray_grid(lo, hi, syn.sun_direction(70.0, 30.0), 300, 240)
    # chunked generation restarts each chunk's origin in float64, so compare against the same construction
    hist = np.zeros(256, np.int64)
    for r0 in range(0, 240, 100):
        nr = min(100, 240 - r0)
        o0 = (grid[0].astype(np.float64) + r0 * grid[2].astype(np.float64)).astype(np.float32)
        import ctypes
        from pyqsm_b200 import _lib
        rays = torch.empty(300 * nr, 6, dtype=torch.float32, device="cuda")
        F3 = lambda x: (ctypes.c_float * 3)(*[float(y) for y in x])
        _lib.check(_lib.load().qsmrt_gen_parallel_rays(ctypes.c_void_p(rays.data_ptr()), 300, nr, F3(o0), F3(grid[1]), F3(grid[2]), F3(grid[3]), None))
        torch.cuda.synchronize()
        c = o.count_intersections(rays.cpu().numpy(), 1)
        hist += np.bincount(np.minimum(c, 255), minlength=256)
    assert np.array_equal(res["intersections"].cpu().numpy(), hist)
    assert res["rays"] == 300 * 240 and 0.0 < res["intercepted_fraction"] < 1.0 and res["mean_layers"] > 0


def test_peel_projection_equals_iterated_cast(oracle_mod):
    """methods.md:53-55 "raycasting projection": every layer's triangle set equals an oracle cast against the
    triangles still present; areas are the hit triangles' 3-D and z-flattened areas (ray_casting.py:285-301)."""
    from pyqsm_b200 import RaycastingScene, environment as env, _lib
    import ctypes as C
    v, t = syn.qsm_tree_mesh(seed=6, n_cylinders=12)
    g = RaycastingScene(output_device="cuda")
    g.add_triangles(v, t)
    nu, nv = 160, 140
    res = env.peel_projection(g, direction=(0, 0, -1), grid=(nu, nv))
    layer_of = res["layer_of"].cpu().numpy()
    assert len(res["layers"]) >= 3 and layer_of.shape == (t.shape[0],)
    st = g.stats()
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    grid = syn.parallel_ray_grid(lo, hi, np.array([0, 0, -1.0]), nu, nv)
    rays = torch.empty(nu * nv, 6, dtype=torch.float32, device="cuda")
    F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
    _lib.check(_lib.load().qsmrt_gen_parallel_rays(C.c_void_p(rays.data_ptr()), nu, nv, F3(grid[0]), F3(grid[1]), F3(grid[2]), F3(grid[3]), None))
    torch.cuda.synchronize()
    rays = rays.cpu().numpy()
    alive = np.ones(t.shape[0], bool)
    p0, p1, p2 = (v[t[:, k]].astype(np.float64) for k in range(3))
    ng = np.cross(p1 - p0, p2 - p0)
    a3, ap = 0.5 * np.linalg.norm(ng, axis=1), 0.5 * np.abs(ng[:, 2])
    for k, (cnt, area3, areap) in enumerate(res["layers"]):
        ids = np.nonzero(alive)[0]
        o = oracle_mod.OracleScene()
        o.add_triangles(v, t[ids])
        ref = o.cast_rays(rays, 1)
        hit = np.unique(ids[ref["primitive_ids"][ref["primitive_ids"] != 0xFFFFFFFF]])
        assert np.array_equal(np.nonzero(layer_of == k)[0], hit), f"layer {k}"
        assert cnt == len(hit)
        np.testing.assert_allclose([area3, areap], [a3[hit].sum(), ap[hit].sum()], rtol=1e-5)
        alive[hit] = False
    # the loop stopped because the rays see nothing any more
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t[np.nonzero(alive)[0]]) if alive.any() else None
    if alive.any():
        assert not np.isfinite(o.cast_rays(rays, 1)["t_hit"]).any()
    assert np.array_equal(layer_of == -1, alive)
    np.testing.assert_allclose(res["area_projected"], ap[~alive].sum(), rtol=1e-5)
    assert res["area_projected"] > 1.5 * ap[layer_of == 0].sum()          # overlap is counted separately
