"""Shared checker for tests/golden/kat.json (analytic vectors)."""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-5      # north_star: t_hit and normals within 1e-5 relative


def load_kats():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        return json.load(f)


def _num(x):
    a = np.array([[math.inf if v == "inf" else v for v in row] if isinstance(row, list) else (math.inf if row == "inf" else row)
                  for row in x], dtype=np.float64)
    return a


def _np(x):
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


def check_kat(kat, scene_factory):
    """scene_factory() -> object with the RaycastingScene query API."""
    if "mesh" not in kat:
        return
    s = scene_factory()
    gid = s.add_triangles(np.asarray(kat["mesh"]["v"], np.float32), np.asarray(kat["mesh"]["t"], np.uint32))
    assert gid == 0
    if "mesh2" in kat:
        assert s.add_triangles(np.asarray(kat["mesh2"]["v"], np.float32), np.asarray(kat["mesh2"]["t"], np.uint32)) == 1
    if "points" in kat:
        q = np.asarray(kat["points"], np.float32)
        if "closest" in kat:
            c, exp = s.compute_closest_points(q), kat["closest"]
            np.testing.assert_allclose(_np(c["points"]), np.asarray(exp["points"]), rtol=RTOL, atol=1e-7, err_msg=kat["name"])
            for key in ("geometry_ids", "primitive_ids"):
                assert np.array_equal(_np(c[key]).astype(np.int64), np.asarray(exp[key], np.int64)), (kat["name"], key)
            np.testing.assert_allclose(_np(c["primitive_uvs"]), np.asarray(exp["primitive_uvs"]), rtol=RTOL, atol=1e-6)
            np.testing.assert_allclose(_np(c["primitive_normals"]), np.asarray(exp["primitive_normals"]), rtol=RTOL, atol=1e-7)
        if "distance" in kat:
            np.testing.assert_allclose(_np(s.compute_distance(q)), np.asarray(kat["distance"]), rtol=RTOL, err_msg=kat["name"])
        if "occupancy" in kat:
            assert _np(s.compute_occupancy(q)).tolist() == kat["occupancy"], kat["name"]
        if "signed_distance" in kat:
            np.testing.assert_allclose(_np(s.compute_signed_distance(q)), np.asarray(kat["signed_distance"]), rtol=RTOL, err_msg=kat["name"])
        return
    rays = np.asarray(kat["rays"], np.float32)
    if "shape" in kat:
        # leading-shape preservation: the same rays shaped [..., 6] give results shaped [...] (+ 2 / 3)
        lead = tuple(kat["shape"])
        shaped = rays.reshape(lead + (6,))
        a = s.cast_rays(shaped)
        assert tuple(_np(a["t_hit"]).shape) == lead and tuple(_np(a["geometry_ids"]).shape) == lead
        assert tuple(_np(a["primitive_ids"]).shape) == lead and tuple(_np(a["primitive_uvs"]).shape) == lead + (2,)
        assert tuple(_np(a["primitive_normals"]).shape) == lead + (3,)
        assert tuple(_np(s.count_intersections(shaped)).shape) == lead and tuple(_np(s.test_occlusions(shaped)).shape) == lead
        flat = s.cast_rays(rays)
        for key in ("t_hit", "geometry_ids", "primitive_ids", "primitive_uvs", "primitive_normals"):
            assert np.array_equal(_np(a[key]).reshape(_np(flat[key]).shape), _np(flat[key])), (kat["name"], key)
    for case in kat.get("occlusion_cases", []):
        tfar = math.inf if case["tfar"] == "inf" else case["tfar"]
        o = _np(s.test_occlusions(rays, tnear=case["tnear"], tfar=tfar))
        assert o.tolist() == case["expect"], (kat["name"], case, o.tolist())
    if "cast" in kat:
        ans = s.cast_rays(rays)
        exp = kat["cast"]
        t = _np(ans["t_hit"]).astype(np.float64)
        et = _num(exp["t_hit"])
        assert np.array_equal(np.isfinite(t), np.isfinite(et)), (kat["name"], t, et)
        m = np.isfinite(et)
        np.testing.assert_allclose(t[m], et[m], rtol=RTOL, err_msg=kat["name"])
        assert np.all(np.isposinf(t[~m]))
        for key in ("geometry_ids", "primitive_ids"):
            if key in exp:
                assert np.array_equal(_np(ans[key]).astype(np.int64), np.asarray(exp[key], np.int64)), (kat["name"], key)
        if "primitive_uvs" in exp:
            np.testing.assert_allclose(_np(ans["primitive_uvs"]), np.asarray(exp["primitive_uvs"]), rtol=RTOL, atol=1e-7)
        if "primitive_normals" in exp:
            np.testing.assert_allclose(_np(ans["primitive_normals"]), np.asarray(exp["primitive_normals"]), rtol=RTOL, atol=1e-7)
        # hits are misses-or-valid: ids are INVALID exactly on misses, uv/normals zero there
        miss = ~np.isfinite(t)
        assert np.all(_np(ans["primitive_ids"])[miss] == 0xFFFFFFFF)
        assert np.all(_np(ans["geometry_ids"])[miss] == 0xFFFFFFFF)
        assert np.all(_np(ans["primitive_uvs"])[miss] == 0) and np.all(_np(ans["primitive_normals"])[miss] == 0)
    if "count" in kat:
        c = _np(s.count_intersections(rays))
        assert c.dtype == np.int32
        assert c.tolist() == kat["count"], (kat["name"], c.tolist())
    if "occluded" in kat:
        o = _np(s.test_occlusions(rays))
        assert o.tolist() == kat["occluded"], (kat["name"], o.tolist())
    if "list" in kat:
        l = s.list_intersections(rays)
        assert _np(l["ray_splits"]).tolist() == kat["list"]["ray_splits"], kat["name"]
        np.testing.assert_allclose(_np(l["t_hit"]), np.asarray(kat["list"]["t_hit"]), rtol=RTOL, err_msg=kat["name"])
        k = len(kat["list"]["t_hit"])
        assert _np(l["ray_ids"]).shape == (k,) and _np(l["primitive_uvs"]).shape == (k, 2)
        splits = _np(l["ray_splits"])
        assert np.array_equal(_np(l["ray_ids"]), np.repeat(np.arange(len(rays)), np.diff(splits)))
        for key in ("geometry_ids", "primitive_ids"):
            if key in kat["list"]:
                assert np.array_equal(_np(l[key]).astype(np.int64), np.asarray(kat["list"][key], np.int64)), (kat["name"], key)
        if "primitive_uvs" in kat["list"]:
            np.testing.assert_allclose(_np(l["primitive_uvs"]), np.asarray(kat["list"]["primitive_uvs"]), rtol=RTOL, atol=1e-6, err_msg=kat["name"])
