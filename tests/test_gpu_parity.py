"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through
the C ABI / RaycastingScene against the CPU oracle on identical inputs.
Bit-exact for hit/miss, ids, counts, and (because the arithmetic is shared
op for op) for t/uv/normals too; the north_star tolerance of 1e-5 relative
on t and normals is asserted as the contract."""
import numpy as np
import pytest
import torch

from kat_util import check_kat, load_kats, RTOL
from pyqsm_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
KATS = load_kats()


@pytest.fixture(scope="module")
def RS():
    from pyqsm_b200 import RaycastingScene
    return RaycastingScene


def _np(x):
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def assert_cast_equal(g, o, flags=None, what=""):
    """GPU dict vs oracle dict.  Returns number of flagged mismatching rays."""
    gp, op_ = _np(g["primitive_ids"]).reshape(-1), o["primitive_ids"].reshape(-1)
    gg, og = _np(g["geometry_ids"]).reshape(-1), o["geometry_ids"].reshape(-1)
    bad = (gp != op_) | (gg != og)
    if bad.any():
        assert flags is not None and np.all(flags.reshape(-1)[bad] != 0), \
            f"{what}: {bad.sum()} id mismatches outside the edge-flagged set"
    ok = ~bad
    gt, ot = _np(g["t_hit"]).reshape(-1)[ok], o["t_hit"].reshape(-1)[ok]
    assert np.array_equal(np.isfinite(gt), np.isfinite(ot))
    m = np.isfinite(ot)
    np.testing.assert_allclose(gt[m], ot[m], rtol=RTOL)
    np.testing.assert_allclose(_np(g["primitive_normals"]).reshape(-1, 3)[ok], o["primitive_normals"].reshape(-1, 3)[ok], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(_np(g["primitive_uvs"]).reshape(-1, 2)[ok], o["primitive_uvs"].reshape(-1, 2)[ok], rtol=1e-4, atol=1e-6)
    # the arithmetic is shared op for op, so in fact everything is bit-identical
    assert np.array_equal(gt, ot), f"{what}: t_hit not bit-identical"
    assert np.array_equal(_np(g["primitive_uvs"]).reshape(-1, 2)[ok], o["primitive_uvs"].reshape(-1, 2)[ok])
    assert np.array_equal(_np(g["primitive_normals"]).reshape(-1, 3)[ok], o["primitive_normals"].reshape(-1, 3)[ok])
    return int(bad.sum())


@pytest.mark.parametrize("kat", [k for k in KATS if "mesh" in k], ids=lambda k: k["name"])
def test_golden_vectors_gpu(RS, kat):
    check_kat(kat, RS)


def test_pinhole_kat(RS):
    kat = [k for k in KATS if k["name"] == "kat8_pinhole"][0]
    rays = RS.create_rays_pinhole(**kat["pinhole"])
    assert tuple(rays.shape) == (2, 4, 6) and rays.dtype == torch.float32
    np.testing.assert_allclose(rays[..., 3:].numpy(), np.asarray(kat["expect_dirs"]), rtol=1e-6, atol=1e-7)
    assert np.all(rays[..., :3].numpy() == np.asarray(kat["expect_origin"], np.float32))
    # intrinsic/extrinsic overload gives the same rays
    from pyqsm_b200.raycasting_scene import _fov_to_matrices, _np64
    p = kat["pinhole"]
    K, E = _fov_to_matrices(p["fov_deg"], _np64(p["center"]), _np64(p["eye"]), _np64(p["up"]), p["width_px"], p["height_px"])
    r2 = RS.create_rays_pinhole(K, E, p["width_px"], p["height_px"])
    assert torch.equal(rays, r2)
    # reference call shape: ray_casting.py:271-277 (fov 90, 1280x950)
    r3 = RS.create_rays_pinhole(fov_deg=90, center=[0.5, 0.2, 3.0], eye=[0.5, 0.2, 13.0], up=[0, 1, -1], width_px=1280, height_px=950)
    assert tuple(r3.shape) == (950, 1280, 6)
    ctr = r3[475, 640, 3:].numpy()           # pixel centre (640.5, 475.5): half a pixel off the axis
    assert abs(ctr[0]) < 2e-3 and ctr[2] < 0


@pytest.fixture(scope="module")
def c1(oracle_mod, RS):
    """BASELINE config C1: 50k-triangle cylinder-QSM tree, 1M parallel sun rays."""
    v, t = syn.qsm_tree_mesh(seed=1)
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    g = RS(output_device="cuda")
    g.add_triangles(v, t)
    return v, t, o, g


def test_c1_cast_rays_full(c1):
    v, t, o, g = c1
    assert t.shape[0] == 50000
    grid = syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(45, 135), 1000, 1000)
    rays = torch.from_numpy(syn.materialize_grid(*grid, 1000, 1000)).cuda()
    ans = {k: a.cpu() for k, a in g.cast_rays(rays).items()}
    ref = o.cast_rays(rays.cpu().numpy(), 1)
    nbad = assert_cast_equal(ans, ref, flags=None, what="C1")
    assert nbad == 0
    assert 0.02 < np.isfinite(ref["t_hit"]).mean() < 0.5


def test_c1_count_and_occlusion(c1):
    v, t, o, g = c1
    grid = syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(45, 135), 500, 500)
    rays = syn.materialize_grid(*grid, 500, 500)
    gc = g.count_intersections(torch.from_numpy(rays)).cpu().numpy()
    oc = o.count_intersections(rays, 1)
    assert gc.dtype == np.int32 and np.array_equal(gc, oc)
    assert oc.max() >= 4
    go = g.test_occlusions(torch.from_numpy(rays)).cpu().numpy()
    assert np.array_equal(go, oc > 0)
    go2 = g.test_occlusions(torch.from_numpy(rays), tnear=20.0, tfar=25.0).cpu().numpy()
    assert np.array_equal(go2, o.test_occlusions(rays, 20.0, 25.0, mode=1))


def test_c1_list_intersections(c1):
    v, t, o, g = c1
    rays = syn.random_rays(v.min(0), v.max(0), 20000, seed=5)
    gl = {k: a.cpu().numpy() for k, a in g.list_intersections(rays).items()}
    ol = o.list_intersections(rays, 1)
    assert gl["ray_splits"].dtype == np.int64 and gl["ray_ids"].dtype == np.int64
    assert gl["geometry_ids"].dtype == np.uint32 and gl["primitive_ids"].dtype == np.uint32
    for k in ol:
        assert np.array_equal(gl[k], ol[k]), k
    assert ol["ray_splits"][-1] > 1000


def test_list_and_count_with_the_per_thread_kernels(c1):
    """traversal_variant 1 (the A/B reference, also the path of trees too deep for the shared-memory stack): no stash
    of hit records, every ray with hits is enumerated hit by hit -- the same CSR arrays."""
    v, t, o, g = c1
    rays = syn.random_rays(v.min(0), v.max(0), 4000, seed=6)
    ol = o.list_intersections(rays, 1)
    try:
        g.set_option("traversal_variant", 1)
        gl = {k: a.cpu().numpy() for k, a in g.list_intersections(rays).items()}
        gc = g.count_intersections(torch.from_numpy(rays)).cpu().numpy()
    finally:
        g.set_option("traversal_variant", 2)
    for k in ol:
        assert np.array_equal(gl[k], ol[k]), k
    assert np.array_equal(gc, np.diff(ol["ray_splits"]).astype(np.int32))


def test_list_intersections_many_hits_per_ray(RS, oracle_mod):
    """list_intersections is ONE all-hits traversal whose records wait in a stash sized for 4 hits per ray; a batch
    with more (here: 12 plates, every ray crosses all of them, some along the plates' diagonals) is traversed again
    with exactly the room it needs.  Two geometries, so geometry ids are part of the records."""
    vs, ts = [], []
    for k in range(12):
        z = np.float32(k)
        vs.append(np.array([[0, 0, z], [10, 0, z], [10, 10, z], [0, 10, z]], np.float32))
        ts.append(np.array([[0, 1, 2], [0, 2, 3]], np.uint32) + np.uint32(4 * k))
    v, t = np.concatenate(vs), np.concatenate(ts)
    o, g = oracle_mod.OracleScene(), RS()
    for s in (o, g):
        assert s.add_triangles(v[:24], t[:12]) == 0
        assert s.add_triangles(v[24:], t[12:] - np.uint32(24)) == 1
    grid = syn.parallel_ray_grid((0, 0, 0), (10, 10, 11), (0, 0, -1), 400, 300, 0.0)
    rays = syn.materialize_grid(*grid, 400, 300)
    rays[::7, 1] = rays[::7, 0]                                   # some rays exactly through the shared diagonals
    gl = {k: a.numpy() for k, a in g.list_intersections(rays).items()}
    ol = o.list_intersections(rays, 0)
    assert ol["ray_splits"][-1] > 4 * len(rays) + 65536           # more records than the first stash holds
    assert np.array_equal(np.diff(ol["ray_splits"]), np.full(len(rays), 12))
    for k in ol:
        assert np.array_equal(gl[k], ol[k]), k
    assert np.array_equal(g.count_intersections(rays).numpy(), np.full(len(rays), 12, np.int32))


def test_c1_vs_brute_force_subsample(c1):
    """GPU BVH traversal against the oracle's brute force (ground truth)."""
    v, t, o, g = c1
    rays = np.concatenate([syn.random_rays(v.min(0), v.max(0), 6000, seed=11),
                           syn.materialize_grid(*syn.parallel_ray_grid(v.min(0), v.max(0), (0, 0, -1), 64, 64), 64, 64)])
    ans = {k: a.cpu() for k, a in g.cast_rays(rays).items()}
    ref = o.cast_rays(rays, 0)
    flags = o.edge_flags(rays, mode=1)
    assert_cast_equal(ans, ref, flags, "C1-brute")
    assert np.array_equal(g.count_intersections(rays).cpu().numpy(), o.count_intersections(rays, 0))


def test_builder_matches_oracle(c1):
    """LBVH builder parity: sorted Morton keys, ordering, Karras topology and
    refit boxes are bit-identical to the oracle's canonical LBVH."""
    import ctypes as C
    from pyqsm_b200 import _lib
    from pyqsm_b200 import RaycastingScene
    v, t, o, _ = c1
    g = RaycastingScene()
    g.set_option("keep_binary_nodes", 1)                     # the product build keeps no binary node array
    g.set_option("split_max", 1)                             # the canonical tree has one leaf per triangle: no sliver splitting
    g.add_triangles(v, t)
    g.commit()
    n = t.shape[0]
    keys = np.empty(n, np.uint64)
    order = np.empty(n, np.uint32)
    nodes = np.empty((2 * n - 1, 8), np.float32)
    _lib.check(g._L.qsmrt_debug_get_build(g._h, keys.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p),
                                          nodes.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(keys, o.sorted_keys())
    assert np.all(np.diff(keys.astype(np.uint64)) >= 0) or np.all(keys[1:] >= keys[:-1])
    assert np.array_equal(order, o.sorted_order())
    olo, ohi, oleft, oright = o.nodes()
    ni = nodes.view(np.int32)
    left, right = ni[: n - 1, 3].copy(), ni[: n - 1, 7].copy()
    # unified index space -> oracle's ~leaf encoding
    conv = lambda c: np.where(c >= n - 1, ~(c - (n - 1)), c)
    assert np.array_equal(conv(left), oleft) and np.array_equal(conv(right), oright)
    assert np.array_equal(nodes[: n - 1, 0:3], olo) and np.array_equal(nodes[: n - 1, 4:7], ohi)
    st = g.stats()
    assert st["num_triangles"] == n and abs(st["box_pad"] - o.box_pad()) == 0
    lo, hi = o.scene_bounds()
    assert np.array_equal(np.asarray(st["scene_lo"], np.float32), lo) and np.array_equal(np.asarray(st["scene_hi"], np.float32), hi)


def _builder_topology(RS, oracle_mod, v, t, keep, opts=None):
    """(left, right, lo, hi) of the GPU's binary tree next to the oracle's, plus a cast through the product nodes."""
    import ctypes as C
    from pyqsm_b200 import _lib
    L = _lib.load()
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    o.commit()
    g = RS()
    g.set_option("keep_binary_nodes", 1 if keep else 0)
    g.set_option("split_max", 1)                             # compare with the oracle's one-leaf-per-triangle tree
    for name, value in (opts or {}).items():
        g.set_option(name, value)
    g.add_triangles(v, t)
    g.commit()
    n = t.shape[0]
    if keep:
        keys, order = np.empty(n, np.uint64), np.empty(n, np.uint32)
        nodes = np.empty((2 * n - 1, 8), np.float32)
        _lib.check(L.qsmrt_debug_get_build(g._h, keys.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p),
                                           nodes.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(keys, o.sorted_keys()) and np.array_equal(order, o.sorted_order())
        if n > 1:
            olo, ohi, oleft, oright = o.nodes()
            ni = nodes.view(np.int32)
            conv = lambda c: np.where(c >= n - 1, ~(c - (n - 1)), c)
            assert np.array_equal(conv(ni[: n - 1, 3]), oleft) and np.array_equal(conv(ni[: n - 1, 7]), oright)
            assert np.array_equal(nodes[: n - 1, 0:3], olo) and np.array_equal(nodes[: n - 1, 4:7], ohi)
    return o, g


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 256, 257, 511, 513, 1000, 4097, 70001])
@pytest.mark.parametrize("kind", ["random", "duplicates", "clusters", "coincident"])
def test_builder_block_boundaries_and_equal_keys(RS, oracle_mod, n, kind):
    """The fused bottom-up builder joins subtrees through shuffles (32 leaves), shared memory (256 leaves) and
    global flags: sizes around those boundaries, and runs of identical Morton keys (index tie-break, the deepest
    trees), must give the oracle's top-down Karras tree bit for bit -- and the product nodes the same hits."""
    rng = np.random.default_rng(n * 7 + len(kind))
    if kind == "random":
        c = rng.uniform(-3, 3, size=(n, 3))
    elif kind == "duplicates":                       # every triangle 1..40 times: long runs of equal keys
        c = np.repeat(rng.uniform(-3, 3, size=(n // 20 + 1, 3)), 40, axis=0)[rng.permutation((n // 20 + 1) * 40)[:n]]
    elif kind == "coincident":                       # one triangle n times (plus two far corners that span the scene):
        c = np.tile(rng.uniform(-1, 1, size=(1, 3)), (n, 1))      # a run the 40-bit sort cannot order -> full-sort fallback
        if n > 2:
            c[:2] = [[-3, -3, -3], [3, 3, 3]]
    else:                                            # a few tight clusters: deep prefixes, unbalanced tree
        c = rng.uniform(-3, 3, size=(5, 3))[rng.integers(0, 5, n)] + rng.normal(0, 1e-4, size=(n, 3))
    d = np.tile(np.array([[0.05, 0, 0], [0, 0.05, 0], [0, 0, 0.0]]), (n, 1, 1)) if kind in ("duplicates", "coincident") else rng.normal(0, 0.05, size=(n, 3, 3))
    v = (c[:, None, :] + d).reshape(-1, 3).astype(np.float32)
    t = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    rays = np.concatenate([syn.random_rays((-3, -3, -3), (3, 3, 3), 1500, seed=n),
                           np.concatenate([v.reshape(n, 3, 3).mean(1)[: min(n, 500)] + np.float32([0, 0, 8.0]), np.tile([0, 0, -1.0], (min(n, 500), 1))], axis=1).astype(np.float32)])
    for keep in (True, False):
        o, g = _builder_topology(RS, oracle_mod, v, t, keep)
        mode = 0 if n <= 4097 else 1
        ref = o.cast_rays(rays, mode)
        assert_cast_equal(g.cast_rays(rays), ref, o.edge_flags(rays, mode=mode), f"{kind}{n}")
        assert np.array_equal(g.count_intersections(rays).numpy(), o.count_intersections(rays, mode))
        assert np.isfinite(ref["t_hit"][1500:]).sum() >= min(n, 500) // 2      # rays over the centroids
        if kind == "coincident":                     # > 64 keys equal in their top 40 bits: eight passes instead of 5 + fix-up
            assert g.stats()["full_sort"] == (1 if n >= 255 else 0) or 33 < n < 255
        else:
            assert g.stats()["full_sort"] == 0


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_soup_vs_brute(RS, oracle_mod, seed):
    """Random triangle soups with slivers, duplicates, degenerates, shared edges, two geometries."""
    rng = np.random.default_rng(seed)
    nv, nt = 300, 1500
    v = rng.uniform(-2, 2, size=(nv, 3)).astype(np.float32)
    t = rng.integers(0, nv, size=(nt, 3)).astype(np.uint32)
    t[0] = (5, 5, 9)
    t[1] = t[2]
    v2, t2 = syn.box_mesh((-1, -1, -1), (1, 1, 1))
    o = oracle_mod.OracleScene()
    g = RS()
    for s in (o, g):
        assert s.add_triangles(v, t) == 0
        assert s.add_triangles(v2, t2) == 1
    rays = np.concatenate([syn.random_rays((-2, -2, -2), (2, 2, 2), 8000, seed=seed + 100),
                           syn.materialize_grid(*syn.parallel_ray_grid((-2, -2, -2), (2, 2, 2), (0, 0, -1), 50, 50, 0.0), 50, 50)])
    ref = o.cast_rays(rays, 0)
    flags = o.edge_flags(rays, mode=0)
    ans = g.cast_rays(rays)
    assert_cast_equal(ans, ref, flags, f"soup{seed}")
    gc = g.count_intersections(rays).numpy()
    oc = o.count_intersections(rays, 0)
    bad = gc != oc
    assert np.all(flags[bad] != 0)
    assert oc.max() > 32, "soup should overflow the per-lane hit set"     # exercises the exact slow path (k_count_fix)
    assert bad.sum() == 0
    gl = {k: a.numpy() for k, a in g.list_intersections(rays).items()}
    ol = o.list_intersections(rays, 0)
    for k in ol:
        assert np.array_equal(gl[k], ol[k]), k


def test_edge_cases(RS):
    s = RS()
    rays = np.array([[0, 0, 0, 0, 0, 1], [1, 1, 1, 1, 0, 0]], np.float32)
    a = s.cast_rays(rays)                                         # empty scene: all miss
    assert torch.all(torch.isinf(a["t_hit"])) and a["primitive_ids"].numpy().tolist() == [0xFFFFFFFF] * 2
    assert s.count_intersections(rays).tolist() == [0, 0]
    assert s.test_occlusions(rays).tolist() == [False, False]
    l = s.list_intersections(rays)
    assert l["ray_splits"].tolist() == [0, 0, 0] and l["t_hit"].numel() == 0
    v, t = syn.box_mesh()
    s.add_triangles(v, t)
    bad = np.array([[np.nan, 0.5, -1, 0, 0, 1], [0.5, 0.5, -1, 0, np.nan, 1], [0.5, 0.5, -1, 0, 0, 0]], np.float32)
    a = s.cast_rays(bad)
    assert torch.all(torch.isinf(a["t_hit"])) and s.count_intersections(bad).tolist() == [0, 0, 0]
    # zero rays, and leading shapes are preserved
    z = s.cast_rays(np.zeros((0, 6), np.float32))
    assert z["t_hit"].shape == (0,) and z["primitive_uvs"].shape == (0, 2)
    r3 = np.zeros((3, 5, 6), np.float32)
    r3[..., 0:2] = 0.4
    r3[..., 2] = -1
    r3[..., 5] = 1
    a = s.cast_rays(r3)
    assert a["t_hit"].shape == (3, 5) and a["primitive_normals"].shape == (3, 5, 3) and torch.all(a["t_hit"] == 1)
    assert s.count_intersections(r3).shape == (3, 5) and torch.all(s.count_intersections(r3) == 2)
    # Open3D error behaviour: dtype / shape
    with pytest.raises(RuntimeError):
        s.cast_rays(np.zeros((4, 6), np.float64))
    with pytest.raises(RuntimeError):
        s.cast_rays(np.zeros((4, 5), np.float32))
    with pytest.raises(RuntimeError):
        s.add_triangles(v, t.astype(np.int64))
    with pytest.raises(RuntimeError):
        s.add_triangles(v, np.array([[0, 1, 99]], np.uint32))
    # single triangle scene
    s1 = RS()
    s1.add_triangles(np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    assert s1.cast_rays(np.array([[0.2, 0.1, 1, 0, 0, -1]], np.float32))["t_hit"].item() == 1.0


def test_host_path_equals_device_path(c1, RS):
    """qsmrt_cast_rays_host (chunked, 3 streams) == qsmrt_cast_rays on resident buffers."""
    v, t, o, g = c1
    h = RS()                       # CPU outputs
    h.add_triangles(v, t)
    grid = syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(30, 10), 1500, 1500)   # 2.25M rays: 3 chunks
    rays = syn.materialize_grid(*grid, 1500, 1500)
    a = h.cast_rays(rays)
    b = g.cast_rays(torch.from_numpy(rays).cuda())
    for k in a:
        assert a[k].device.type == "cpu"
        assert torch.equal(a[k], b[k].cpu()), k


def test_reference_call_patterns(RS):
    """The call shapes of pyQSM/viz/ray_casting.py run unchanged against the replacement."""
    from pyqsm_b200.mesh import TriangleMesh
    v, t = syn.qsm_tree_mesh(seed=2, n_cylinders=40)
    mesh = TriangleMesh(v, t)
    rcs = RS
    # raycast_to_pcd, ray_casting.py:315-322
    scene = rcs()
    scene.add_triangles(mesh)
    center = mesh.get_center().numpy()
    cfg = {'fov_deg': 90, 'center': mesh.get_center(), 'eye': [center[0], center[1], center[2] + 10], 'up': [0, 1, -1],
           'width_px': 640, 'height_px': 480}
    rays = rcs.create_rays_pinhole(**cfg)
    ans = scene.cast_rays(rays)
    intersecting_rays = ans['t_hit'].isfinite()
    hits = rays[intersecting_rays]
    points = hits[:, :3] + hits[:, 3:] * ans['t_hit'][intersecting_rays].reshape((-1, 1))
    assert points.shape[1] == 3 and points.shape[0] > 100
    # cast_rays surf_2d branch, ray_casting.py:285-289
    hit_triangle_ids = ans['primitive_ids'][intersecting_rays].numpy()
    hit_tris = np.asarray(t)[hit_triangle_ids]
    assert hit_tris.shape[1] == 3
    flags = scene.mark_hit_primitives(ans).numpy()
    assert np.array_equal(np.nonzero(flags)[0], np.unique(hit_triangle_ids))
    # sparse_cast_w_intersections, ray_casting.py:155-180
    scene = rcs()
    mesh_id = scene.add_triangles(mesh)
    assert mesh_id == 0
    bb_min = mesh.vertex['positions'].min(dim=0).numpy()
    bb_max = mesh.vertex['positions'].max(dim=0).numpy()
    x, y = np.linspace(bb_min, bb_max, num=10)[:, :2].T
    xv, yv = np.meshgrid(x, y)
    orig = np.stack([xv, yv, np.full_like(xv, bb_min[2] - 1)], axis=-1).reshape(-1, 3)
    dest = orig + np.full(orig.shape, (0, 0, 2 + bb_max[2] - bb_min[2]), dtype=np.float32)
    rays = np.concatenate([orig, dest - orig], axis=-1).astype(np.float32)
    lx = scene.list_intersections(rays)
    lx = {k: v_.numpy() for k, v_ in lx.items()}
    vv = mesh.vertex['positions'].numpy()
    tt = mesh.triangle['indices'].numpy()
    tidx = lx['primitive_ids']
    uv = lx['primitive_uvs']
    w = 1 - np.sum(uv, axis=1)
    c_arr = vv[tt[tidx, 1].flatten(), :] * uv[:, 0][:, None] + vv[tt[tidx, 2].flatten(), :] * uv[:, 1][:, None] + \
        vv[tt[tidx, 0].flatten(), :] * w[:, None]
    c_ref = rays[lx['ray_ids']][:, :3] + rays[lx['ray_ids']][:, 3:] * lx['t_hit'][..., None]
    assert len(c_arr) > 0
    np.testing.assert_allclose(c_arr, c_ref, atol=1e-4)
    # get_points_inside_mesh, ray_casting.py:53-71: occupancy of a closed cylinder
    cv, ct = syn.cylinder_mesh(1.0, 2.0)
    scene = rcs()
    scene.add_triangles(TriangleMesh(cv, ct))
    q = np.array([[0, 0, 0], [0.5, 0.2, 0.9], [2, 0, 0], [0, 0, 1.5]], np.float32)
    assert scene.compute_occupancy(q).tolist() == [1.0, 1.0, 0.0, 0.0]


def test_variants_tilings_and_leaf_sizes_agree(RS, oracle_mod):
    """Both traversal kernels (1 per-thread loop, 2 persistent), fp32 and quantised nodes, the 2-D tile
    mapping and every leaf size give bit-identical cast_rays results -- and match the oracle."""
    from pyqsm_b200 import _lib
    L = _lib.load()
    v, t = syn.qsm_tree_mesh(seed=9, n_cylinders=60)
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    grid = syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(35, 200), 301, 203)   # not multiples of the 8x4 tile
    rays = syn.materialize_grid(*grid, 301, 203)
    # signed zeros and axis-parallel directions: -0.0 has a negative reciprocal (near/far plane selection)
    g2 = syn.parallel_ray_grid(v.min(0), v.max(0), np.array([-0.6, -0.0, -0.8]), 301, 203)
    rays[: 301 * 100] = syn.materialize_grid(*g2, 301, 203)[: 301 * 100]
    rays[301 * 100: 301 * 120, 3:] = np.array([0.0, -0.0, -1.0], np.float32)
    rays[301 * 120: 301 * 140, 3:] = np.array([-0.0, 1.0, 0.0], np.float32)
    ref = o.cast_rays(rays, 1)
    rays_img = torch.from_numpy(rays.reshape(203, 301, 6)).cuda()
    g = RS(output_device="cuda")
    g.add_triangles(v, t)
    for leaf_max in (1, 2, 3, 4):
        g.set_option("leaf_max", leaf_max)                           # a changed builder option marks the scene for rebuilding
        g.commit()
        assert g.stats()["leaf_max"] == leaf_max and g.stats()["bvh_height"] >= 10
        assert g.stats()["quantised_nodes"] == 1                     # this mesh qualifies for the 32-byte nodes
        for variant, quant in ((1, 1), (2, 1), (2, 0)):
            g.set_option("traversal_variant", variant)
            g.set_option("quantised_nodes", quant)                   # only the persistent kernel reads them
            for r in (rays_img, rays_img.reshape(-1, 6)):            # 2-D tiles / linear
                ans = {k: a.cpu().reshape((-1,) + tuple(a.shape[r.ndim - 1:])) for k, a in g.cast_rays(r).items()}
                assert_cast_equal(ans, ref, None, f"leaf{leaf_max}/v{variant}/q{quant}")
        g.set_option("traversal_variant", 2)
        g.set_option("quantised_nodes", 1)
        occ = g.test_occlusions(rays_img.reshape(-1, 6)).cpu().numpy()
        assert np.array_equal(occ, np.isfinite(ref["t_hit"]))
    # options live in the scene: a fresh scene has the defaults whatever another scene was set to
    g2 = RS()
    assert g2.get_option("leaf_max") == 2 and g2.get_option("traversal_variant") == 2 and g.get_option("leaf_max") == 4


def test_fetch_counters(RS):
    """set_option('counters', 1): the kernel's own node / triangle fetch counts, per scene."""
    v, t = syn.qsm_tree_mesh(seed=1)
    g = RS(output_device="cuda")
    g.add_triangles(v, t)
    rays = torch.from_numpy(syn.materialize_grid(*syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(45, 135), 400, 400), 400, 400)).cuda()
    g.set_option("counters", 1)
    g.cast_rays(rays)
    c = g.counters()
    assert 1 < c[0] / rays.shape[0] < 200 and 0 < c[1] / rays.shape[0] < 50
    assert c[3] <= 32 * c[2] and c[8] <= 32 * c[7]              # lanes per node / triangle iteration
    other = RS(output_device="cuda")
    assert other.counters() == [0] * 16                          # counters are per scene


def test_closest_points_and_signed_distance(RS, oracle_mod):
    """compute_closest_points / compute_distance / compute_signed_distance (ray_casting.py:250,255)
    bit-identical to the oracle; the mri() call pattern of the reference runs unchanged."""
    from pyqsm_b200.mesh import TriangleMesh
    v, t = syn.qsm_tree_mesh(seed=3, n_cylinders=60)
    v2, t2 = syn.box_mesh((3, 3, 0), (5, 5, 2))
    g, o = RS(), oracle_mod.OracleScene()
    for s in (g, o):
        s.add_triangles(v, t)
        s.add_triangles(v2, t2)
    q = np.random.default_rng(2).uniform(v.min(0) - 1, v.max(0) + 1, size=(50000, 3)).astype(np.float32)
    q[:100] = v[:100]                                   # on-surface queries: distance 0, many ties
    a, r = g.compute_closest_points(q), o.compute_closest_points(q, 1)
    for k in ("points", "geometry_ids", "primitive_ids", "primitive_uvs", "primitive_normals"):
        assert np.array_equal(a[k].numpy(), r[k]), k
    assert np.array_equal(g.compute_distance(q).numpy(), r["distance"])
    sd = g.compute_signed_distance(q).numpy()
    assert np.array_equal(sd, o.compute_signed_distance(q, 1))
    assert (sd < 0).sum() > 50 and np.all(np.abs(sd) == r["distance"])
    # vs brute force on a subsample
    rb = o.compute_closest_points(q[:3000], 0)
    assert np.array_equal(a["primitive_ids"].numpy()[:3000], rb["primitive_ids"])
    # reference pattern, ray_casting.py:241-255 (mri): random points + a 64^3 grid
    v3_, t3_ = syn.box_mesh((3.1, 3.37, 0.23), (5.02, 5.61, 2.11))     # irregular bounds: no ray through an exact edge
    mesh = TriangleMesh(v3_, t3_)
    scene = RS()
    _ = scene.add_triangles(mesh)
    o3 = oracle_mod.OracleScene()
    o3.add_triangles(v3_, t3_)
    min_bound = mesh.vertex.positions.min(0).numpy()
    max_bound = mesh.vertex.positions.max(0).numpy()
    query_points = np.random.uniform(low=min_bound, high=max_bound, size=[256, 3]).astype(np.float32)
    signed_distance = scene.compute_signed_distance(query_points)
    assert signed_distance.shape == (256,) and torch.all(signed_distance <= 0)
    xyz_range = np.linspace(min_bound - 0.5, max_bound + 0.5, num=64)
    query_points = np.stack(np.meshgrid(*xyz_range.T), axis=-1).astype(np.float32)
    signed_distance = scene.compute_signed_distance(query_points)
    assert signed_distance.numpy()[:, :, 10].shape == (64, 64)
    assert np.array_equal(signed_distance.numpy(), o3.compute_signed_distance(query_points, 1))
    inside = np.all((query_points > min_bound) & (query_points < max_bound), axis=-1)
    # single-sample parity (Open3D's default too) is fooled only by rays grazing a box edge exactly
    assert ((signed_distance.numpy() < 0) != inside).mean() < 1e-3
    # empty scene: infinite distance
    e = RS()
    assert torch.all(torch.isinf(e.compute_distance(q[:4])))
    with pytest.raises(RuntimeError):
        g.compute_distance(np.zeros((4, 2), np.float32))


def test_open3d_shim_import_path():
    """SURVEY.md 8b route 2: with the shim on sys.path the reference's own import lines bind the B200 engine
    (ray_casting.py:8,43) and its get_points_inside_mesh pattern (:55-69) runs."""
    import importlib
    import os
    import sys
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyqsm_b200", "open3d_shim")
    sys.path.insert(0, shim)
    try:
        for m in [k for k in sys.modules if k == "open3d" or k.startswith("open3d.")]:
            del sys.modules[m]
        from open3d.t.geometry import RaycastingScene as rcs          # ray_casting.py:8
        import open3d as o3d
        import open3d.core as o3c                                      # ray_casting.py:43
        import pyqsm_b200
        assert rcs is pyqsm_b200.RaycastingScene and rcs.INVALID_ID == 4294967295
        cv, ct = syn.cylinder_mesh(radius=0.5, height=2.0)
        mesh = o3d.t.geometry.TriangleMesh(cv, ct)
        query_pts = np.array([[0, 0, 0], [0.2, 0.1, 0.8], [0.6, 0, 0], [0, 0, 1.2]])
        tpts = o3c.Tensor(query_pts, o3c.float32)                      # ray_casting.py:62
        scene = rcs()
        _ = scene.add_triangles(mesh)
        occ = scene.compute_occupancy(tpts)                            # ray_casting.py:69
        assert occ.numpy().tolist() == [1.0, 1.0, 0.0, 0.0]
    finally:
        sys.path.remove(shim)
        for m in [k for k in sys.modules if k == "open3d" or k.startswith("open3d.")]:
            del sys.modules[m]
        importlib.invalidate_caches()


def test_c2_full_size_properties(RS, oracle_mod):
    """BASELINE config C2 at full size (2M-triangle canopy, 16M parallel sun rays): size-independent
    properties of the domain on every ray, and the oracle on a 1M-ray subsample (every 16th ray)."""
    import ctypes as C
    from pyqsm_b200 import _lib
    L = _lib.load()
    v, t = syn.canopy_mesh(2, 1_000_000)
    assert t.shape[0] == 2_000_000
    g = RS(output_device="cuda")
    g.add_triangles(v, t)
    build_ms = g.commit()
    st = g.stats()
    assert st["num_triangles"] == 2_000_000 and build_ms < 50.0
    nu = nv = 4000
    grid = syn.parallel_ray_grid(np.asarray(st["scene_lo"]), np.asarray(st["scene_hi"]), syn.sun_direction(40, 135), nu, nv)
    F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
    rays = torch.empty(nv, nu, 6, dtype=torch.float32, device="cuda")
    _lib.check(L.qsmrt_gen_parallel_rays(C.c_void_p(rays.data_ptr()), nu, nv, F3(grid[0]), F3(grid[1]), F3(grid[2]), F3(grid[3]), None))
    a = g.cast_rays(rays)                                   # [4000,4000,6]: 8x4 tiles, persistent kernel, quantised nodes
    hit = torch.isfinite(a["t_hit"])
    assert 0.2 < hit.float().mean().item() < 0.8
    # hit <=> valid ids; misses carry INVALID_ID and zero uv / normal; normals are unit length
    inv = torch.tensor(0xFFFFFFFF, dtype=torch.int64, device="cuda")
    pid = a["primitive_ids"].to(torch.int64)
    assert torch.equal(pid != inv, hit) and torch.equal(a["geometry_ids"].to(torch.int64) != inv, hit)
    assert int(pid[hit].max()) < 2_000_000
    assert torch.all(a["primitive_uvs"][~hit] == 0) and torch.all(a["primitive_normals"][~hit] == 0)
    nlen = torch.linalg.norm(a["primitive_normals"][hit], dim=-1)
    assert torch.all((nlen - 1).abs() < 1e-5)
    uv = a["primitive_uvs"][hit]
    assert torch.all(uv >= 0) and torch.all(uv.sum(-1) <= 1 + 1e-6)
    # occlusion == hit; count > 0 <=> hit (test_occlusions / count_intersections, other kernels modes)
    flat = rays.reshape(-1, 6)
    assert torch.equal(g.test_occlusions(flat).reshape(nv, nu), hit)
    cnt = g.count_intersections(flat[: 4_000_000])
    assert torch.equal(cnt > 0, hit.reshape(-1)[: 4_000_000])
    # scaling d by 4 (exact in fp32) scales t by 1/4 and changes nothing else; linear and tiled mappings agree
    r2 = flat.clone()
    r2[:, 3:] *= 4.0
    b = g.cast_rays(r2)
    assert torch.equal(b["primitive_ids"].reshape(nv, nu), a["primitive_ids"])
    assert torch.equal(b["t_hit"].reshape(nv, nu)[hit] * 4.0, a["t_hit"][hit])
    assert torch.equal(b["primitive_uvs"].reshape(nv, nu, 2), a["primitive_uvs"])
    # hit point from barycentrics == o + t d (ray_casting.py:172-180 convention), on a sample
    idx = torch.nonzero(hit.reshape(-1))[:: 997, 0]
    tri = torch.from_numpy(t.astype(np.int64)).cuda()[pid.reshape(-1)[idx]]
    vv = torch.from_numpy(v).cuda().double()
    uvs = a["primitive_uvs"].reshape(-1, 2)[idx].double()
    p = vv[tri[:, 1]] * uvs[:, :1] + vv[tri[:, 2]] * uvs[:, 1:] + vv[tri[:, 0]] * (1 - uvs.sum(1, keepdim=True))
    q = flat[idx, :3].double() + flat[idx, 3:].double() * a["t_hit"].reshape(-1)[idx].double()[:, None]
    assert (p - q).abs().max().item() < 5e-5
    # the oracle on every 16th ray: bit-identical
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    sub = flat[::16].cpu().numpy()
    ref = o.cast_rays(sub, 1)
    ans = {k: x.reshape((-1,) + tuple(x.shape[2:]))[::16].cpu() for k, x in a.items()}
    assert assert_cast_equal(ans, ref, None, "C2 full size") == 0


def test_add_cylinders_on_device(RS, oracle_mod):
    """QSM cylinder records -> mesh on the GPU: same topology and vertices as Open3D's
    create_cylinder + get_shape() placement built on the host, and rays see the same scene."""
    from pyqsm_b200.synthetic import _rotation_to
    rng = np.random.default_rng(8)
    n = 40
    centers = rng.uniform(-3, 3, size=(n, 3))
    axes = rng.normal(size=(n, 3))
    axes[0] = (0, 0, 1); axes[1] = (0, 0, -1); axes[2] = (1, 0, 0)
    radii = rng.uniform(0.05, 0.4, size=n)
    heights = rng.uniform(0.3, 2.0, size=n)
    g = RS()
    gid = g.add_cylinders(centers, axes, radii, heights)
    assert gid == 0
    v_dev, t_dev = g.geometry(gid)
    cv, ct = syn.cylinder_mesh(1.0, 1.0, 20, 4)
    vs, ts = [], []
    for k in range(n):
        R = _rotation_to(axes[k])
        vk = (cv.astype(np.float64) * np.array([radii[k], radii[k], heights[k]])) @ R.T + centers[k]
        vs.append(vk)
        ts.append(ct + np.uint32(k * cv.shape[0]))
    v_ref, t_ref = np.concatenate(vs).astype(np.float32), np.concatenate(ts)
    assert v_dev.shape == (n * 102, 3) and t_dev.shape == (n * 200, 3)
    assert np.array_equal(t_dev.numpy(), t_ref)
    np.testing.assert_allclose(v_dev.numpy(), v_ref, atol=2e-6 * 8)
    # the device-generated geometry traces like the same mesh handed to the oracle
    o = oracle_mod.OracleScene()
    o.add_triangles(v_dev.numpy(), t_dev.numpy())
    rays = syn.random_rays(v_ref.min(0), v_ref.max(0), 20000, seed=4)
    a, r = g.cast_rays(rays), o.cast_rays(rays, 1)
    assert_cast_equal(a, r, None, "cylinders")
    assert np.array_equal(g.count_intersections(rays).numpy(), o.count_intersections(rays, 1))
    inside = g.compute_occupancy(centers.astype(np.float32)).numpy()
    assert inside.mean() > 0.9                                           # cylinder centres are inside their cylinders


from hypothesis import given, settings  # noqa: E402
from conftest import HYPOTHESIS_DERANDOMIZE  # noqa: E402
from test_oracle import _mesh_and_rays  # noqa: E402


@settings(max_examples=25, deadline=None, derandomize=HYPOTHESIS_DERANDOMIZE)
@given(_mesh_and_rays())
def test_property_gpu_equals_brute_force(case):
    """Adversarial small meshes (shared edges, slivers, duplicates, degenerates, three scales; rays through
    vertices / edge midpoints / axis-parallel): every GPU query equals the oracle's brute force, bit for bit."""
    import oracle
    from pyqsm_b200 import RaycastingScene
    v, t, rays = case
    o, g = oracle.OracleScene(), RaycastingScene()
    o.add_triangles(v, t)
    g.add_triangles(v, t)
    ref = o.cast_rays(rays, 0)
    assert assert_cast_equal(g.cast_rays(rays), ref, None, "property") == 0
    assert np.array_equal(g.count_intersections(rays).numpy(), o.count_intersections(rays, 0))
    assert np.array_equal(g.test_occlusions(rays).numpy(), np.isfinite(ref["t_hit"]))
    gl, ol = g.list_intersections(rays), o.list_intersections(rays, 0)
    for k in ol:
        assert np.array_equal(gl[k].numpy(), ol[k]), k
    gp, op_ = g.compute_closest_points(rays[:, :3].copy()), o.compute_closest_points(rays[:, :3], 0)
    for k in gp:        # a zero-area triangle can be the closest primitive: its normal is 0/0 = NaN on both sides
        assert np.array_equal(gp[k].numpy(), op_[k], equal_nan=op_[k].dtype.kind == "f"), k


def test_degenerate_triangle_as_closest_primitive(RS, oracle_mod):
    """The case hypothesis found: a point-triangle [2, 2, 2] sitting on a mesh vertex comes out 1 ulp closer than the
    real triangles around that vertex (a different branch of the region test), so it is the closest primitive; ids,
    point, distance and uv agree with the oracle and the normal of the zero-area triangle is NaN (0 / 0) in both."""
    z = np.float32(62.5)
    v = np.array([[0, 0, -z], [250, 0, 0], [500, 0, z], [0, 250, z], [250, 250, z], [500, 250, 0], [0, 500, z], [250, 500, z],
                  [500, 500, -z]], np.float32)
    t = np.array([[0, 1, 4], [0, 4, 3], [1, 2, 5], [1, 5, 4], [3, 4, 7], [3, 7, 6], [4, 5, 8], [4, 8, 7], [0, 1, 4], [0, 4, 3], [1, 2, 5],
                  [0, 0, 1], [2, 2, 2]], np.uint32)
    q = np.array([[107.0203, -950.06494, 1633.9882], [520, -10, 80], [100, 100, 300], [-50, -60, -200]], np.float32)
    o, g = oracle_mod.OracleScene(), RS()
    o.add_triangles(v, t)
    g.add_triangles(v, t)
    a, r = g.compute_closest_points(q), o.compute_closest_points(q, 0)
    assert r["primitive_ids"][0] == 12 and np.all(np.isnan(r["primitive_normals"][0]))
    for k in a:
        assert np.array_equal(a[k].numpy(), r[k], equal_nan=r[k].dtype.kind == "f"), k
    assert np.array_equal(g.compute_distance(q).numpy(), r["distance"])


def test_far_origins_and_offset_scenes(RS, oracle_mod):
    """Conservative boxes at any distance: the leaf padding (2^-17 of the scene, 3 cells for the quantised nodes)
    covers origins within a few tens of scene sizes; beyond that every slab interval is widened by 2^-20 of its own
    t (trace_persistent.cuh SLAB_NEAR / SLAB_FAR), which grows with the distance like the rounding of the slab
    constants does.  Origins 20, 1000 and 10000 scene diagonals away, a tree translated 300 m from the coordinate
    origin (LiDAR plot offsets), both node formats, both traversal kernels: every hit brute force finds is found."""
    v, t = syn.qsm_tree_mesh(seed=11, n_cylinders=40)
    diag = float(np.linalg.norm(v.max(0) - v.min(0)))
    for shift in (np.zeros(3, np.float32), np.array([300.0, -250.0, 40.0], np.float32)):
        vs = (v + shift).astype(np.float32)
        o = oracle_mod.OracleScene()
        o.add_triangles(vs, t)
        near = syn.random_rays(vs.min(0), vs.max(0), 6000, seed=3)
        sets = [near]
        for k in (20.0, 1.0e3, 1.0e4):
            far = near.copy()
            far[:, :3] -= far[:, 3:] * np.float32(k * diag)            # same lines, origins k diagonals back
            sets.append(far)
        rays = np.concatenate(sets)
        ref = o.cast_rays(rays, 0)                                     # brute force
        for k in range(1, 4):
            assert np.isfinite(ref["t_hit"][6000 * k: 6000 * (k + 1)]).sum() > 300
        cnt = o.count_intersections(rays, 0)
        for quant, variant in ((1, 2), (0, 2), (1, 1)):
            g = RS()
            g.set_option("quantised_nodes", quant)
            g.set_option("traversal_variant", variant)
            g.add_triangles(vs, t)
            assert_cast_equal(g.cast_rays(rays), ref, None, f"far/shift{shift[0]}/q{quant}/v{variant}")
            assert np.array_equal(g.count_intersections(rays).numpy(), cnt)


@pytest.mark.parametrize("cap", [1, 7, 100])
def test_builder_climb_list_overflow(RS, oracle_mod, cap):
    """Open subtrees that do not fit the climb list finish their climb inside the hierarchy kernel: same tree."""
    from pyqsm_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(cap)
    n = 20011
    c = rng.uniform(-3, 3, size=(n, 3))
    v = (c[:, None, :] + rng.normal(0, 0.05, size=(n, 3, 3))).reshape(-1, 3).astype(np.float32)
    t = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    rays = syn.random_rays((-3, -3, -3), (3, 3, 3), 3000, seed=cap)
    for keep in (True, False):
        o, g = _builder_topology(RS, oracle_mod, v, t, keep, {"climb_capacity": cap})
        assert_cast_equal(g.cast_rays(rays), o.cast_rays(rays, 1), o.edge_flags(rays, mode=1), f"climbcap{cap}")


def test_scene_churn_reuses_device_blocks(RS, oracle_mod):
    """The reference makes a new scene in every function; libqsmrt recycles the device blocks of dead scenes.
    Scenes of equal size built one after the other must not see each other's data, and the cache can be dropped."""
    import pyqsm_b200
    rays = syn.random_rays((-3, -3, -3), (3, 3, 3), 4000, seed=3)
    for rep in range(6):
        rng = np.random.default_rng(100 + rep)
        n = 3000 if rep % 2 == 0 else 2999
        v = rng.uniform(-3, 3, size=(3 * n, 3)).astype(np.float32)
        v[1::3] = v[0::3] + rng.normal(0, 0.3, size=(n, 3)).astype(np.float32)
        v[2::3] = v[0::3] + rng.normal(0, 0.3, size=(n, 3)).astype(np.float32)
        t = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
        o = oracle_mod.OracleScene(); o.add_triangles(v, t)
        g = RS(); g.add_triangles(v, t)
        assert_cast_equal(g.cast_rays(rays), o.cast_rays(rays, 0), o.edge_flags(rays, mode=0), f"churn{rep}")
        assert np.array_equal(g.count_intersections(rays).numpy(), o.count_intersections(rays, 0))
        del g
        if rep == 3:
            pyqsm_b200.empty_cache()


def test_occupancy_majority_vote(RS):
    """compute_occupancy / compute_signed_distance with nsamples > 1 (Open3D: odd, majority of nsamples rays): on a
    closed box every direction agrees with the analytic inside test; on a box with one face removed the single
    (1,1,1) ray leaks through the hole for some points while the majority of 5 directions does not."""
    v, t = syn.box_mesh((0, 0, 0), (1, 1, 1))
    g = RS()
    g.add_triangles(v, t)
    rng = np.random.default_rng(7)
    q = rng.uniform(-0.5, 1.5, size=(4000, 3)).astype(np.float32)
    inside = np.all((q > 0) & (q < 1), axis=1)
    for ns in (1, 3, 5):
        assert np.array_equal(g.compute_occupancy(q, nsamples=ns).numpy() > 0, inside)
        sd = g.compute_signed_distance(q, nsamples=ns).numpy()
        assert np.array_equal(sd < 0, inside)
        np.testing.assert_array_equal(np.abs(sd), g.compute_distance(q).numpy())
    with pytest.raises(RuntimeError):
        g.compute_occupancy(q, nsamples=2)
    # open box: drop the two triangles of the +x face (x == 1)
    keep = ~np.all(v[t.astype(np.int64)][:, :, 0] == 1.0, axis=1)
    h = RS()
    h.add_triangles(v, t[keep])
    qi = rng.uniform(0.05, 0.95, size=(2000, 3)).astype(np.float32)
    one = h.compute_occupancy(qi, nsamples=1).numpy() > 0
    five = h.compute_occupancy(qi, nsamples=5).numpy() > 0
    assert one.mean() < 0.9 and five.mean() > one.mean()


def test_closest_points_warp_and_thread_kernels_agree(RS, oracle_mod):
    """Small closest-point batches run one warp per query, large ones one thread per query: same answers, and the
    oracle's, on the C1 tree (points inside, near and far outside the mesh)."""
    from pyqsm_b200 import _lib
    L = _lib.load()
    v, t = syn.qsm_tree_mesh(seed=1, n_cylinders=60)
    g = RS()
    g.add_triangles(v, t)
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    rng = np.random.default_rng(11)
    lo, hi = v.min(0), v.max(0)
    q = np.concatenate([rng.uniform(lo, hi, size=(3000, 3)), rng.uniform(lo - 30, hi + 30, size=(1500, 3)),
                        v[rng.integers(0, len(v), 500)] + rng.normal(0, 1e-3, size=(500, 3))]).astype(np.float32)
    a = {k: x.numpy() for k, x in g.compute_closest_points(q).items()}
    g.set_option("cp_warp_max", 0)
    b = {k: x.numpy() for k, x in g.compute_closest_points(q).items()}
    g.set_option("cp_warp_max", 16384)
    ref = o.compute_closest_points(q, 1)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
        if k in ref:
            assert np.array_equal(a[k], ref[k]), k
    assert np.array_equal(g.compute_distance(q).numpy(), ref["distance"]) if "distance" in ref else True
