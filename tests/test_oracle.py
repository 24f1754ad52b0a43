"""CPU tests of the oracle (no GPU): golden vectors, brute force vs canonical
LBVH, an independent float64 intersector, and the domain invariances of
SURVEY.md section 8c."""
import numpy as np
import pytest

from kat_util import check_kat, load_kats
from pyqsm_b200 import synthetic as syn

KATS = load_kats()


@pytest.fixture(scope="module")
def orc(oracle_mod):
    return oracle_mod


class _ModeScene:
    """OracleScene pinned to one search mode, with the RaycastingScene method names."""

    def __init__(self, orc, mode):
        self.s, self.mode = orc.OracleScene(), mode

    def add_triangles(self, v, t):
        return self.s.add_triangles(v, t)

    def cast_rays(self, r):
        return self.s.cast_rays(r, self.mode)

    def count_intersections(self, r):
        return self.s.count_intersections(r, self.mode)

    def test_occlusions(self, r, tnear=0.0, tfar=float("inf")):
        return self.s.test_occlusions(r, tnear, tfar, mode=self.mode)

    def list_intersections(self, r):
        return self.s.list_intersections(r, self.mode)

    def compute_closest_points(self, q):
        return self.s.compute_closest_points(q, self.mode)

    def compute_distance(self, q):
        return self.s.compute_distance(q, self.mode)

    def compute_occupancy(self, q):
        return self.s.compute_occupancy(q, self.mode)

    def compute_signed_distance(self, q):
        return self.s.compute_signed_distance(q, self.mode)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("kat", [k for k in KATS if "mesh" in k], ids=lambda k: k["name"])
def test_golden_vectors(orc, kat, mode):
    check_kat(kat, lambda: _ModeScene(orc, mode))


def _exact_closest(v, t, rays):
    """Independent float64 ray/triangle intersector (plain barycentric solve)."""
    v = v.astype(np.float64)
    p0, p1, p2 = v[t[:, 0]], v[t[:, 1]], v[t[:, 2]]
    out_t = np.full(len(rays), np.inf)
    out_p = np.full(len(rays), -1, np.int64)
    margin = np.full(len(rays), np.inf)
    for i, r in enumerate(rays.astype(np.float64)):
        o, d = r[:3], r[3:]
        e1, e2 = p1 - p0, p2 - p0
        n = np.cross(e1, e2)
        den = n @ d
        with np.errstate(divide="ignore", invalid="ignore"):
            tt = ((p0 - o) @ n.T).diagonal() / den if False else np.einsum("ij,ij->i", p0 - o, n) / den
            q = o + tt[:, None] * d
            # barycentrics of q
            a = np.einsum("ij,ij->i", np.cross(e1, q - p0), n)
            b = np.einsum("ij,ij->i", np.cross(q - p0, e2), n)
            nn = np.einsum("ij,ij->i", n, n)
            vv, uu = a / nn, b / nn
        ww = 1 - uu - vv
        inside = (den != 0) & (tt > 0) & (uu >= 0) & (vv >= 0) & (ww >= 0)
        m = np.minimum(np.minimum(uu, vv), ww)
        near = (den != 0) & (tt > 0) & (np.abs(m) < 1e-5)
        if near.any():
            margin[i] = 0.0
        if inside.any():
            k = np.where(inside)[0]
            j = k[np.argmin(tt[k])]
            out_t[i], out_p[i] = tt[j], j
            srt = np.sort(tt[k])
            if len(srt) > 1 and (srt[1] - srt[0]) < 1e-5 * max(1.0, srt[0]):
                margin[i] = 0.0
    return out_t, out_p, margin


def test_oracle_vs_exact_float64(orc):
    """The fp32 Embree-formulation oracle agrees with an exact-arithmetic
    intersector away from edges (hit/miss and primitive identical, t to 1e-5)."""
    v, t = syn.qsm_tree_mesh(seed=7, n_cylinders=6)
    lo, hi = v.min(0), v.max(0)
    rays = syn.random_rays(lo, hi, 400, seed=3)
    grid = syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(50, 20), 20, 20), 20, 20)
    rays = np.concatenate([rays, grid])
    s = orc.OracleScene()
    s.add_triangles(v, t)
    a = s.cast_rays(rays, 0)
    et, ep, margin = _exact_closest(v, t, rays)
    ok = margin > 0
    assert ok.sum() > 0.9 * len(rays)
    hit = np.isfinite(et)
    assert np.array_equal(np.isfinite(a["t_hit"])[ok], hit[ok])
    m = ok & hit
    assert m.sum() > 20
    assert np.array_equal(a["primitive_ids"][m].astype(np.int64), ep[m])
    np.testing.assert_allclose(a["t_hit"][m], et[m], rtol=1e-5)


def _exact_point_triangle(p, a, b, c):
    """float64 closest point of triangle abc to p by a formulation unlike Ericson's region walk: the plane
    projection if it falls inside, else the nearest of the three edge segments."""
    n = np.cross(b - a, c - a)
    nn = n @ n
    cands = []
    if nn > 0:
        q = p - n * ((p - a) @ n) / nn
        w = [np.cross(b - a, q - a) @ n, np.cross(c - b, q - b) @ n, np.cross(a - c, q - c) @ n]
        if min(w) >= 0:
            cands.append(q)
    for u, v_ in ((a, b), (b, c), (c, a)):
        d = v_ - u
        dd = d @ d
        s_ = 0.0 if dd == 0 else min(1.0, max(0.0, ((p - u) @ d) / dd))
        cands.append(u + s_ * d)
    return min(cands, key=lambda x: (x - p) @ (x - p))


def test_oracle_closest_point_vs_exact_float64(orc):
    """The oracle's fp32 closest-point query (Ericson regions, as in the Embree tutorial Open3D uses) against an
    independent float64 formulation, brute force over all triangles: distance to 1e-5, the same triangle unless a
    second one is as close, and closest point = (1-u-v) v0 + u v1 + v v2 with the returned uv."""
    v, t = syn.qsm_tree_mesh(seed=5, n_cylinders=2)
    rng = np.random.default_rng(2)
    lo, hi = v.min(0), v.max(0)
    q = np.concatenate([rng.uniform(lo - 1, hi + 1, size=(40, 3)), v[rng.integers(0, len(v), 15)] + rng.normal(0, 0.02, size=(15, 3))]).astype(np.float32)
    s = orc.OracleScene()
    s.add_triangles(v, t)
    for mode in (0, 1):
        a = s.compute_closest_points(q, mode)
        V = v.astype(np.float64)
        for i, p in enumerate(q.astype(np.float64)):
            d2 = np.array([((_exact_point_triangle(p, V[k[0]], V[k[1]], V[k[2]]) - p) ** 2).sum() for k in t])
            best = np.sqrt(d2.min())
            np.testing.assert_allclose(np.linalg.norm(a["points"][i].astype(np.float64) - p), best, rtol=1e-5, atol=1e-6)
            pid = int(a["primitive_ids"][i])
            assert np.sqrt(d2[pid]) <= best * (1 + 1e-5) + 1e-6
            k = t[pid]
            u, w = a["primitive_uvs"][i].astype(np.float64)
            rec = (1 - u - w) * V[k[0]] + u * V[k[1]] + w * V[k[2]]
            np.testing.assert_allclose(rec, a["points"][i], rtol=1e-5, atol=1e-5)
        assert np.array_equal(a["geometry_ids"], np.zeros(len(q), np.uint32))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_brute_equals_bvh(orc, seed):
    rng = np.random.default_rng(seed)
    v, t = syn.qsm_tree_mesh(seed=seed + 10, n_cylinders=20)
    # add slivers, duplicates and a degenerate triangle
    extra_v = rng.uniform(-3, 3, size=(30, 3)).astype(np.float32)
    extra_t = rng.integers(0, 30, size=(40, 3)).astype(np.uint32)
    extra_t[0] = (1, 1, 2)                       # degenerate
    extra_t[1] = extra_t[2]                      # duplicate
    s = orc.OracleScene()
    s.add_triangles(v, t)
    s.add_triangles(extra_v, extra_t)
    lo, hi = v.min(0), v.max(0)
    rays = np.concatenate([syn.random_rays(lo, hi, 3000, seed=seed),
                           syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, (0, 0, -1), 40, 40), 40, 40)])
    a, b = s.cast_rays(rays, 0), s.cast_rays(rays, 1)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(s.count_intersections(rays, 0), s.count_intersections(rays, 1))
    # mode 2: the binned-SAH tree of the tree-quality measurement answers identically (and visits no more nodes)
    lb_nodes = (s.cast_rays(rays, 1), s.last_counters)[1][0]
    c = s.cast_rays(rays, 2)
    assert s.last_counters[0] <= 1.1 * lb_nodes
    for k in a:
        assert np.array_equal(a[k], c[k]), k
    assert np.array_equal(s.count_intersections(rays, 0), s.count_intersections(rays, 2))
    assert np.array_equal(s.test_occlusions(rays, mode=0), s.test_occlusions(rays, mode=1))
    assert np.array_equal(s.test_occlusions(rays, 0.5, 2.0, mode=0), s.test_occlusions(rays, 0.5, 2.0, mode=1))
    la, lb = s.list_intersections(rays, 0), s.list_intersections(rays, 1)
    for k in la:
        assert np.array_equal(la[k], lb[k]), k
    assert np.array_equal(s.edge_flags(rays, mode=0), s.edge_flags(rays, mode=1))


def test_invariances(orc):
    """Scaling d by k scales t by 1/k, ids/uv/normals unchanged; count>0 <=> finite t;
    sum of split diffs = K; occlusion == (count > 0) on the full interval."""
    v, t = syn.qsm_tree_mesh(seed=3, n_cylinders=30)
    s = orc.OracleScene()
    s.add_triangles(v, t)
    rays = syn.random_rays(v.min(0), v.max(0), 5000, seed=9)
    a = s.cast_rays(rays)
    r2 = rays.copy()
    r2[:, 3:] *= 4.0                                  # power of two: exact in fp32
    b = s.cast_rays(r2)
    assert np.array_equal(a["primitive_ids"], b["primitive_ids"])
    assert np.array_equal(a["geometry_ids"], b["geometry_ids"])
    hit = np.isfinite(a["t_hit"])
    assert np.array_equal(a["t_hit"][hit], b["t_hit"][hit] * 4.0)
    assert np.array_equal(a["primitive_uvs"], b["primitive_uvs"])
    assert np.array_equal(a["primitive_normals"], b["primitive_normals"])
    c = s.count_intersections(rays)
    assert np.array_equal(c > 0, hit)
    assert np.array_equal(s.test_occlusions(rays), hit)
    l = s.list_intersections(rays)
    assert l["ray_splits"][-1] == len(l["t_hit"]) == c.sum()
    first = l["ray_splits"][:-1][hit]
    assert np.array_equal(l["t_hit"][first], a["t_hit"][hit])          # sorted by t: first is the closest
    assert np.array_equal(l["primitive_ids"][first], a["primitive_ids"][hit])
    nrm = a["primitive_normals"][hit]
    np.testing.assert_allclose(np.linalg.norm(nrm, axis=1), 1.0, rtol=1e-6)


def test_uv_convention_reconstructs_hit_point(orc):
    """ray_casting.py:172-180: p = u*v1 + v*v2 + (1-u-v)*v0 equals o + t*d."""
    v, t = syn.qsm_tree_mesh(seed=4, n_cylinders=15)
    s = orc.OracleScene()
    s.add_triangles(v, t)
    rays = syn.random_rays(v.min(0), v.max(0), 4000, seed=2)
    a = s.cast_rays(rays)
    hit = np.isfinite(a["t_hit"])
    tri = t[a["primitive_ids"][hit]]
    uv = a["primitive_uvs"][hit].astype(np.float64)
    w = 1 - uv.sum(1)
    p = v[tri[:, 1]] * uv[:, :1] + v[tri[:, 2]] * uv[:, 1:] + v[tri[:, 0]] * w[:, None]
    q = rays[hit, :3].astype(np.float64) + rays[hit, 3:] * a["t_hit"][hit][:, None]
    np.testing.assert_allclose(p, q, atol=2e-5)


def test_empty_and_nan(orc):
    s = orc.OracleScene()
    rays = np.array([[0, 0, 0, 0, 0, 1]], np.float32)
    a = s.cast_rays(rays)
    assert np.isposinf(a["t_hit"][0]) and a["primitive_ids"][0] == 0xFFFFFFFF
    assert s.count_intersections(rays)[0] == 0
    v, t = syn.box_mesh()
    s.add_triangles(v, t)
    bad = np.array([[np.nan, 0.5, -1, 0, 0, 1], [0.5, 0.5, -1, 0, np.nan, 1], [0.5, 0.5, -1, 0, 0, 0]], np.float32)
    a = s.cast_rays(bad)
    assert np.all(np.isposinf(a["t_hit"])) and np.all(s.count_intersections(bad) == 0)
    with pytest.raises(RuntimeError):
        s.add_triangles(v, np.array([[0, 1, 99]], np.uint32))


def test_multi_geometry_ids_and_ties(orc):
    """Two coincident squares in different geometries: equal t -> lowest geometry id wins
    cast_rays; count_intersections counts per geometry."""
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    s = orc.OracleScene()
    assert s.add_triangles(v, t) == 0
    assert s.add_triangles(v, t) == 1
    rays = np.array([[0.7, 0.2, 1, 0, 0, -1], [0.5, 0.5, 1, 0, 0, -1]], np.float32)
    for mode in (0, 1):
        a = s.cast_rays(rays, mode)
        assert a["geometry_ids"].tolist() == [0, 0] and a["primitive_ids"].tolist() == [0, 0]
        assert s.count_intersections(rays, mode).tolist() == [2, 2]
        assert s.edge_flags(rays, mode=mode).tolist() == [2, 3]


def test_closest_points_oracle(orc):
    """Closest-point query: analytic answers on the unit box, brute force == LBVH, and the
    result is a true minimum (an independent float64 point/triangle distance)."""
    v, t = syn.box_mesh()
    s = orc.OracleScene()
    s.add_triangles(v, t)
    q = np.array([[0.5, 0.5, 2], [0.5, 0.5, 0.5], [2, 2, 2], [-1, 0.5, 0.5], [0.25, 0.3, 0.9]], np.float32)
    for mode in (0, 1):
        r = s.compute_closest_points(q, mode)
        np.testing.assert_allclose(r["distance"], [1.0, 0.5, 3 ** 0.5, 1.0, 0.1], rtol=1e-6)
        np.testing.assert_allclose(r["points"], [[0.5, 0.5, 1], [0.5, 0.5, 0], [1, 1, 1], [0, 0.5, 0.5], [0.25, 0.3, 1]], atol=1e-6)
        np.testing.assert_allclose(s.compute_signed_distance(q, mode), [1.0, -0.5, 3 ** 0.5, 1.0, -0.1], rtol=1e-6)
        assert s.compute_occupancy(q, mode).tolist() == [0, 1, 0, 0, 1]
    v, t = syn.qsm_tree_mesh(seed=3, n_cylinders=25)
    s = orc.OracleScene()
    s.add_triangles(v, t)
    q = np.random.default_rng(0).uniform(v.min(0) - 1, v.max(0) + 1, size=(3000, 3)).astype(np.float32)
    a, b = s.compute_closest_points(q, 0), s.compute_closest_points(q, 1)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # uv convention reconstructs the closest point; the point lies on the reported triangle
    tri = t[a["primitive_ids"]]
    uv = a["primitive_uvs"].astype(np.float64)
    p = v[tri[:, 1]] * uv[:, :1] + v[tri[:, 2]] * uv[:, 1:] + v[tri[:, 0]] * (1 - uv.sum(1))[:, None]
    np.testing.assert_allclose(p, a["points"], atol=2e-6)
    np.testing.assert_allclose(np.linalg.norm(a["points"].astype(np.float64) - q, axis=1), a["distance"], rtol=1e-5, atol=1e-6)
    # no sampled surface point is closer than the reported distance
    rng = np.random.default_rng(1)
    w = rng.dirichlet([1, 1, 1], size=t.shape[0])
    samples = (v[t[:, 0]] * w[:, :1] + v[t[:, 1]] * w[:, 1:2] + v[t[:, 2]] * w[:, 2:]).astype(np.float64)
    for i in range(0, 3000, 150):
        assert np.min(np.linalg.norm(samples - q[i], axis=1)) >= a["distance"][i] * (1 - 1e-5)


# ---- property tests (hypothesis): the canonical LBVH is not semantics -- brute force is ----------------
from hypothesis import given, settings, strategies as st  # noqa: E402
from conftest import HYPOTHESIS_DERANDOMIZE  # noqa: E402


@st.composite
def _mesh_and_rays(draw):
    """Small adversarial meshes: a grid patch with shared edges, optional slivers, duplicates and
    degenerate triangles; rays that aim at vertices / edge midpoints as well as random ones."""
    n = draw(st.integers(2, 5))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(n + 1, dtype=np.float64), np.arange(n + 1, dtype=np.float64))
    zs = rng.integers(-1, 2, size=xs.shape) * 0.25 * draw(st.sampled_from([0.0, 1.0]))
    v = np.stack([xs.ravel(), ys.ravel(), zs.ravel()], 1)
    tri = []
    for j in range(n):
        for i in range(n):
            a, b, c, d = j * (n + 1) + i, j * (n + 1) + i + 1, (j + 1) * (n + 1) + i + 1, (j + 1) * (n + 1) + i
            tri += [(a, b, c), (a, c, d)]
    tri = np.asarray(tri, np.uint32)
    if draw(st.booleans()):
        tri = np.concatenate([tri, tri[: draw(st.integers(1, 4))]])               # exact duplicates
    if draw(st.booleans()):
        tri = np.concatenate([tri, np.array([[0, 0, 1], [2, 2, 2]], np.uint32)])      # degenerate
    if draw(st.booleans()):
        v = np.concatenate([v, [[0.5, 0.5, 1.0], [0.5 + 1e-6, 0.5, 1.0], [2.5, 2.5, 1.0]]])
        tri = np.concatenate([tri, np.array([[len(v) - 3, len(v) - 2, len(v) - 1]], np.uint32)])   # sliver
    scale = draw(st.sampled_from([1.0, 1e-3, 250.0]))
    v = (v * scale).astype(np.float32)
    k = 40
    targets = np.concatenate([v[rng.integers(0, len(v), k)],                                    # through vertices
                              0.5 * (v[tri[rng.integers(0, len(tri), k), 0]] + v[tri[rng.integers(0, len(tri), k), 1]]),
                              rng.uniform(v.min(0), v.max(0) + 1e-6, size=(k, 3))]).astype(np.float64)
    origins = targets + rng.normal(size=targets.shape) * scale * 3 + np.array([0, 0, 4.0 * scale])
    axis = rng.integers(0, 3, len(targets))
    d = targets - origins
    snap = rng.random(len(targets)) < 0.3                                         # some axis-parallel rays
    d[snap] = 0
    d[snap, axis[snap]] = -scale
    rays = np.concatenate([origins, d], 1).astype(np.float32)
    return v, tri, rays


@settings(max_examples=40, deadline=None, derandomize=HYPOTHESIS_DERANDOMIZE)
@given(_mesh_and_rays())
def test_property_brute_equals_bvh(case):
    import oracle
    v, t, rays = case
    s = oracle.OracleScene()
    s.add_triangles(v, t)
    a, b = s.cast_rays(rays, 0), s.cast_rays(rays, 1)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    c0, c1 = s.count_intersections(rays, 0), s.count_intersections(rays, 1)
    assert np.array_equal(c0, c1)
    assert np.array_equal(c0 > 0, np.isfinite(a["t_hit"]))
    assert np.array_equal(s.test_occlusions(rays, mode=0), s.test_occlusions(rays, mode=1))
    l0, l1 = s.list_intersections(rays, 0), s.list_intersections(rays, 1)
    for k in l0:
        assert np.array_equal(l0[k], l1[k]), k
    assert np.array_equal(np.diff(l0["ray_splits"]), c0)
    q = rays[:, :3]
    p0, p1 = s.compute_closest_points(q, 0), s.compute_closest_points(q, 1)
    for k in p0:        # a zero-area triangle can be the closest primitive: its normal is 0/0 = NaN on both sides
        assert np.array_equal(p0[k], p1[k], equal_nan=p0[k].dtype.kind == "f"), k
