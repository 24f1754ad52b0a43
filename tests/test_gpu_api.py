"""GPU tests of the boundary features around the hot path (run with -m gpu):
per-scene options, lazy / selected cast_rays outputs, the host pipes of
count_intersections / test_occlusions, scene files, per-vertex exposure, the
Open3D mesh overload with Int64 indices, two threads with their own scenes."""
import threading

import numpy as np
import pytest
import torch

from pyqsm_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def RS():
    from pyqsm_b200 import RaycastingScene
    return RaycastingScene


@pytest.fixture(scope="module")
def tree(oracle_mod):
    v, t = syn.qsm_tree_mesh(seed=4, n_cylinders=80)
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    rays = syn.materialize_grid(*syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(40, 70), 600, 500), 600, 500)
    return v, t, o, rays


def test_mesh_overload_casts_int64_indices(RS, tree):
    """Open3D's add_triangles(mesh) casts positions to Float32 and indices to UInt32 (tensor meshes made by
    from_legacy / create_cylinder carry Int64 indices); only the two-tensor overload is strict.  Every reference
    call site uses the mesh overload (ray_casting.py:66,156,219,242,276,317)."""
    v, t, o, rays = tree

    class _Mesh:                                   # duck-typed open3d.t.geometry.TriangleMesh
        def __init__(self, v, t):
            self.vertex = {"positions": torch.from_numpy(v.astype(np.float64))}
            self.triangle = {"indices": torch.from_numpy(t.astype(np.int64))}

    s = RS()
    assert s.add_triangles(_Mesh(v, t)) == 0
    a = s.cast_rays(rays[:20000])
    ref = o.cast_rays(rays[:20000], 1)
    assert np.array_equal(a["primitive_ids"].numpy(), ref["primitive_ids"]) and np.array_equal(a["t_hit"].numpy(), ref["t_hit"])
    with pytest.raises(RuntimeError):              # the explicit two-tensor form stays strict, like Open3D
        RS().add_triangles(v, t.astype(np.int64))
    bad = _Mesh(v, t)
    bad.triangle["indices"] = bad.triangle["indices"].clone()
    bad.triangle["indices"][0, 0] = -1
    with pytest.raises(RuntimeError):
        RS().add_triangles(bad)


def test_cast_result_lazy_selected_and_all(RS, tree):
    """Default: five keys, t_hit / primitive_ids on the host at once, the rest fetched on first access; "all": plain
    dict; a tuple: only those keys.  Same numbers whichever way they travel."""
    from pyqsm_b200.raycasting_scene import CastResult, CAST_KEYS
    v, t, o, rays = tree
    s = RS()
    s.add_triangles(v, t)
    ref = o.cast_rays(rays, 1)
    for src in (rays, torch.from_numpy(rays).cuda()):            # host pipe / resident rays
        a = s.cast_rays(src)
        assert isinstance(a, CastResult) and set(a.pending()) == {"geometry_ids", "primitive_uvs", "primitive_normals"}
        assert list(a.keys()) == list(CAST_KEYS) and len(a) == 5 and "primitive_uvs" in a
        assert a["t_hit"].device.type == "cpu" and a["t_hit"].shape == (rays.shape[0],)
        for k in CAST_KEYS:
            assert a[k].device.type == "cpu" and np.array_equal(a[k].numpy(), ref[k]), k
        assert a.pending() == ()
        full = s.cast_rays(src, outputs="all")
        assert type(full) is dict and all(np.array_equal(full[k].numpy(), ref[k]) for k in CAST_KEYS)
        two = s.cast_rays(src, outputs=("t_hit", "primitive_ids"))
        assert sorted(two) == ["primitive_ids", "t_hit"] and np.array_equal(two["primitive_ids"].numpy(), ref["primitive_ids"])
    img = s.cast_rays(rays.reshape(500, 600, 6))
    assert img["primitive_normals"].shape == (500, 600, 3) and img["t_hit"].shape == (500, 600)
    assert dict(img.items()).keys() == set(CAST_KEYS)
    with pytest.raises(RuntimeError):
        s.cast_rays(rays[:10], outputs=("t_hit", "nope"))


def test_count_and_occlusion_host_pipes(RS, tree):
    """count_intersections / test_occlusions with host rays go through the same chunked three-stream pipe as
    cast_rays: identical to the resident-buffer calls and to the oracle."""
    v, t, o, rays = tree
    big = np.concatenate([rays] * 5)[: (1 << 20) + 12345]         # > one chunk, ragged tail
    s = RS()
    s.add_triangles(v, t)
    dev = torch.from_numpy(big).cuda()
    cnt_dev, occ_dev = s.count_intersections(dev), s.test_occlusions(dev, tnear=0.5, tfar=9.0)
    # stages of 64k rays with short stages at both ends (8k, 16k, 32k ... 32k, 16k, 8k), equal stages, the default size
    for chunk, ramp in ((65536, 1), (65536, 0), (0, 1)):
        s.set_option("host_chunk", chunk)
        s.set_option("host_ramp", ramp)
        cnt = s.count_intersections(big)
        occ = s.test_occlusions(big, tnear=0.5, tfar=9.0)
        assert cnt.device.type == "cpu" and torch.equal(cnt, cnt_dev), (chunk, ramp)
        assert occ.dtype == torch.bool and torch.equal(occ, occ_dev), (chunk, ramp)
        full = s.cast_rays(big[:400_000], outputs="all")
        assert all(np.array_equal(full[k].numpy()[: rays.shape[0]], o.cast_rays(rays, 1)[k]) for k in ("t_hit", "primitive_ids", "primitive_normals"))
    with pytest.raises(RuntimeError):
        s.set_option("host_chunk", 100)
    n = rays.shape[0]
    assert np.array_equal(cnt.numpy()[:n], o.count_intersections(rays, 1))
    assert np.array_equal(occ.numpy()[:n], o.test_occlusions(rays, 0.5, 9.0, 1))


@pytest.mark.parametrize("with_bvh", [True, False])
def test_scene_file_round_trip(RS, tree, tmp_path, with_bvh):
    """save -> load gives the same scene: geometry ids, stats, and bit-identical query results, whether the file
    carries the LBVH or the loader rebuilds it (the build is deterministic)."""
    v, t, o, rays = tree
    s = RS()
    s.add_triangles(v, t)
    v2, t2 = syn.box_mesh((-1, -1, 0), (1, 1, 2))
    assert s.add_triangles(v2, t2) == 1                           # two geometries
    s.set_option("leaf_max", 3)
    a = s.cast_rays(rays, outputs="all")
    cnt = s.count_intersections(rays)
    path = tmp_path / "scene.qsmrt"
    s.save(path, with_bvh=with_bvh)
    st = s.stats()
    del s
    z = RS.load(path)
    if with_bvh:
        assert z.stats()["num_bvh_nodes"] == st["num_bvh_nodes"] and z.stats()["bvh_height"] == st["bvh_height"]
    b = z.cast_rays(rays, outputs="all")
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(cnt, z.count_intersections(rays))
    zs = z.stats()
    assert zs["num_triangles"] == st["num_triangles"] and zs["num_geometries"] == 2 and zs["leaf_max"] == 3
    assert zs["num_bvh_nodes"] == st["num_bvh_nodes"]
    gv, gt = z.geometry(1)
    assert np.array_equal(gv.numpy(), v2) and np.array_equal(gt.numpy(), t2)
    assert z.add_triangles(v2 + np.float32(5.0), t2) == 2         # a loaded scene stays editable
    assert z.stats()["num_triangles"] == st["num_triangles"] + 12
    with pytest.raises(RuntimeError):
        bad = tmp_path / "bad.qsmrt"
        bad.write_bytes(b"not a scene file" * 40)
        RS.load(bad)


def test_scene_file_damage_is_an_error_not_a_fault(RS, tree, tmp_path):
    """A scene file cut short, or whose BVH points outside its own arrays, raises RuntimeError at load (every
    reference the traversal would follow is range-checked on the GPU); the process stays usable."""
    v, t, o, rays = tree
    s = RS()
    s.add_triangles(v, t)
    path = tmp_path / "scene.qsmrt"
    s.save(path, with_bvh=True)
    st = s.stats()
    raw = bytearray(path.read_bytes())
    lv, nn, q = st["num_references"], max(st["num_references"] - 1, 1), 1 if st["quantised_nodes"] else 0
    tnodes_at = len(raw) - (lv * (8 + 4 + 48) + nn * 32 * q + nn * 64)       # keys, order, records, quantised twin, nodes
    cut = tmp_path / "cut.qsmrt"
    cut.write_bytes(bytes(raw[: len(raw) - 1000]))
    with pytest.raises(RuntimeError, match="truncated"):
        RS.load(cut)
    wild = bytearray(raw)
    wild[tnodes_at + 48: tnodes_at + 56] = np.asarray([0x7FFFFFF0, 0x7FFFFFF0], np.int32).tobytes()      # the root's child references
    bad = tmp_path / "wild.qsmrt"
    bad.write_bytes(bytes(wild))
    with pytest.raises(RuntimeError, match="out of range"):
        RS.load(bad)
    z = RS.load(path)                                               # the undamaged file still loads and answers
    assert torch.equal(z.cast_rays(rays)["t_hit"], s.cast_rays(rays)["t_hit"])


def test_vertex_exposure_and_hit_vertices(RS, tree):
    """Per-vertex results as the reference derives them (ray_casting.py:287-292): hit_tris = triangles[prim_ids],
    hit_vert_ids = np.unique(hit_tris).  mark_hit_primitives(vertices=True) gives that set as a mask,
    vertex_exposure the counts (a vertex inherits the hits of the triangles it is a corner of)."""
    from pyqsm_b200 import environment as env
    v, t, o, rays = tree
    s = RS(output_device="cuda")
    s.add_triangles(v, t)
    ans = s.cast_rays(torch.from_numpy(rays).cuda())
    ref = o.cast_rays(rays, 1)
    prim = ref["primitive_ids"][np.isfinite(ref["t_hit"])]
    hit_tris = t[prim]
    hit_vert_ids = np.unique(hit_tris)
    tri_mask, vert_mask = s.mark_hit_primitives(ans, vertices=True)
    assert np.array_equal(np.flatnonzero(tri_mask.cpu().numpy()), np.unique(prim))
    assert np.array_equal(np.flatnonzero(vert_mask.cpu().numpy()), hit_vert_ids)
    tri_counts = np.bincount(prim, minlength=t.shape[0]).astype(np.int32)
    vc = s.vertex_exposure(torch.from_numpy(tri_counts).cuda()).cpu().numpy()
    want = np.zeros(v.shape[0], np.int64)
    np.add.at(want, t.reshape(-1), np.repeat(tri_counts, 3))
    assert np.array_equal(vc, want)
    # the fused sun driver reports both
    angles = [(40.0, 70.0), (65.0, 250.0)]
    r = env.sun_exposure(s, angles, grid=(300, 200), per_vertex=True)
    assert r["vertex_counts"].shape == (v.shape[0],) and int(r["vertex_counts"].sum()) == 3 * int(r["counts"].sum())
    assert np.array_equal(r["vertex_counts"].cpu().numpy() > 0, np.isin(np.arange(v.shape[0]), t[r["counts"].cpu().numpy() > 0]))


def test_two_threads_with_their_own_options(RS, tree):
    """No process-global tuning state: two threads build and query scenes with different leaf sizes / traversal
    kernels at the same time and each sees its own options and the right answers."""
    v, t, o, rays = tree
    ref = o.cast_rays(rays[:50000], 1)
    errors, seen = [], {}

    def work(name, leaf_max, variant):
        try:
            for rep in range(4):
                s = RS()
                s.set_option("leaf_max", leaf_max)
                s.set_option("traversal_variant", variant)
                s.add_triangles(v, t)
                a = s.cast_rays(rays[:50000], outputs=("t_hit", "primitive_ids"))
                assert s.stats()["leaf_max"] == leaf_max and s.get_option("traversal_variant") == variant
                assert np.array_equal(a["primitive_ids"].numpy(), ref["primitive_ids"])
                assert np.array_equal(a["t_hit"].numpy(), ref["t_hit"])
                seen[name] = s.stats()["num_bvh_nodes"]
        except Exception as e:                      # noqa: BLE001
            errors.append((name, repr(e)))

    th = [threading.Thread(target=work, args=("a", 1, 2)), threading.Thread(target=work, args=("b", 4, 1))]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors
    assert seen["a"] > seen["b"]                    # one triangle per leaf makes more nodes than four


def test_current_device_is_left_alone(RS, tree):
    """Every ABI entry restores the caller's CUDA device (a scene on another GPU must not move torch's)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    v, t, o, rays = tree
    torch.cuda.set_device(0)
    s = RS(device="cuda:1")
    s.add_triangles(v, t)
    a = s.cast_rays(rays[:1000])
    del s
    assert torch.cuda.current_device() == 0 and a["t_hit"].shape == (1000,)


def test_sliver_splitting_changes_the_tree_not_the_answers(RS, oracle_mod):
    """Long thin triangles (create_cylinder sides) enter the LBVH as several references with tight slab boxes
    (QSMRT_OPT_SPLIT_MAX / _ASPECT).  Every query answers exactly as without splitting -- duplicates of a triangle
    are one hit -- and meshes without slivers are left alone."""
    from pyqsm_b200 import environment as env
    v, t = syn.qsm_tree_mesh(seed=5, n_cylinders=60)
    o = oracle_mod.OracleScene()
    o.add_triangles(v, t)
    rays = np.concatenate([syn.materialize_grid(*syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(45, 135), 400, 300), 400, 300),
                           syn.random_rays(v.min(0), v.max(0), 20000, seed=2)])
    ref = o.cast_rays(rays, 0)                                      # brute force: no tree at all
    refc = o.count_intersections(rays, 0)
    refl = o.list_intersections(rays, 0)
    q = syn.random_rays(v.min(0), v.max(0), 3000, seed=4)[:, :3].copy()
    refq = o.compute_closest_points(q, 0)
    seen = {}
    for split_max, leaf_max in ((8, 2), (1, 2), (16, 4), (3, 1)):
        s = RS()
        s.set_option("split_max", split_max)
        s.set_option("leaf_max", leaf_max)
        s.add_triangles(v, t)
        a = s.cast_rays(rays, outputs="all")
        for k in ref:
            assert np.array_equal(a[k].numpy(), ref[k]), (split_max, k)
        assert np.array_equal(s.count_intersections(rays).numpy(), refc)
        assert np.array_equal(s.test_occlusions(rays, 0.25, 7.0).numpy(), o.test_occlusions(rays, 0.25, 7.0, 0))
        l = s.list_intersections(rays)
        for k in refl:
            assert np.array_equal(l[k].numpy(), refl[k]), (split_max, k)
        c = s.compute_closest_points(q)
        for k in ("points", "geometry_ids", "primitive_ids", "primitive_uvs", "primitive_normals"):
            assert np.array_equal(c[k].numpy(), refq[k]), (split_max, k)
        st = s.stats()
        seen[split_max] = st["num_references"]
        assert st["num_triangles"] == t.shape[0] and (st["num_references"] > t.shape[0]) == (split_max > 1)
        if split_max == 8:                                          # the peel driver keeps its flags per triangle, not per reference
            sd = RS(output_device="cuda")
            sd.add_triangles(v, t)
            sp = RS(output_device="cuda")
            sp.set_option("split_max", 1)
            sp.add_triangles(v, t)
            p1, p2 = env.peel_projection(sd, grid=(300, 300), max_layers=6), env.peel_projection(sp, grid=(300, 300), max_layers=6)
            assert torch.equal(p1["layer_of"], p2["layer_of"]) and [x[0] for x in p1["layers"]] == [x[0] for x in p2["layers"]]
            np.testing.assert_allclose(np.asarray(p1["layers"]), np.asarray(p2["layers"]), rtol=1e-12)    # areas: sums of double atomics
            path = "/tmp/qsmrt_split_scene.bin"
            s.save(path)
            z = RS.load(path)
            assert z.stats()["num_references"] == st["num_references"]
            assert torch.equal(z.cast_rays(rays, outputs=("primitive_ids",))["primitive_ids"], a["primitive_ids"])
    assert seen[16] >= seen[8] > seen[3] > seen[1] == t.shape[0]
    vc, tc = syn.canopy_mesh(7, 20000)                              # compact leaf triangles: nothing to split
    sc = RS()
    sc.add_triangles(vc, tc)
    sc.commit()
    assert sc.stats()["num_references"] == tc.shape[0]
