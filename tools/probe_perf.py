"""A/B probe of the traversal kernels for one build of libqsmrt (QSMRT_LIB selects the .so): C2 cast_rays on three
solar angles, the fused sun sweep, the sky Monte-Carlo, count / list on the canopy, and the C1 tree with both
traversal kernels.  Prints one line per measurement; not product code.
    QSMRT_LIB=build/variants/libqsmrt_x.so python tools/probe_perf.py [c2 sun sky cnt c1]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, environment as env, _lib
L = _lib.load()
tag = os.path.basename(os.environ.get("QSMRT_LIB", "libqsmrt.so"))
want = sys.argv[1:] or ["c2", "sun", "sky", "cnt", "c1"]
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def gtime(f, reps=3):
    best = 1e30
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
def grid_rays(s, d, nu, nv):
    st = s.stats(); g = syn.parallel_ray_grid(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), d, nu, nv)
    r = torch.empty(nu * nv, 6, dtype=torch.float32, device="cuda")
    _lib.check(L.qsmrt_gen_parallel_rays(P(r), nu, nv, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None)); return r
def outs(n):
    return [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
if set(want) & {"c2", "sun", "sky", "cnt"}:
    v, t = syn.canopy_mesh(2, 1_000_000)
    s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit()
if "c2" in want:
    G = 4000; o = outs(G * G); tot = 0.0
    for el, az in ((10, 225), (50, 0), (80, 135)):
        r = grid_rays(s, syn.sun_direction(el, az), G, G)
        ms = gtime(lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), G, G, *[P(x) for x in o], None))); tot += ms
        print(f"{tag} C2 cast el{el} az{az}: {ms:.3f} ms {G*G/ms/1e3:.0f} Mr/s", flush=True)
    ms2 = gtime(lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), G, G, P(o[0]), None, P(o[2]), None, None, None)))
    print(f"{tag} C2 cast (3 angles) mean {3*G*G/tot/1e3:.0f} Mr/s | t_hit+prim only, last angle: {G*G/ms2/1e3:.0f} Mr/s", flush=True)
    del r, o
if "sun" in want:
    sweep = syn.hemisphere_sweep()[::4]
    env.sun_exposure(s, sweep[:2], grid=(4000, 4000)); torch.cuda.synchronize()
    t0 = time.perf_counter(); env.sun_exposure(s, sweep, grid=(4000, 4000)); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{tag} fused sun sweep 16 angles: {len(sweep)*16e6/dt/1e6:.0f} Mr/s", flush=True)
if "sky" in want:
    tri = t.reshape(-1, 2, 3)[:, 0]; p0 = v[tri[:, 0]]; nrm = np.cross(v[tri[:, 1]] - p0, v[tri[:, 2]] - p0); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts, nd = torch.from_numpy(p0[::4].copy()).cuda(), torch.from_numpy(nrm[::4].astype(np.float32)).cuda()
    for rf in (12, 8, 16):
        s.set_option("refill", rf)
        env.sky_gap_fraction(s, pts[:10000], nd[:10000], n_dirs=100); torch.cuda.synchronize()
        t0 = time.perf_counter(); env.sky_gap_fraction(s, pts, nd, n_dirs=1000); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{tag} sky 250k pts x 1000 dirs refill {rf}: {pts.shape[0]*1e3/dt/1e6:.0f} Mr/s", flush=True)
    s.set_option("refill", 12)
if "cnt" in want:
    r = grid_rays(s, syn.sun_direction(70, 0), 2000, 2000); n = r.shape[0]; cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = gtime(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(r), n, P(cnt), None)))
    print(f"{tag} canopy count_intersections 4M rays: {n/ms/1e3:.0f} Mr/s (mean count {cnt.float().mean().item():.2f})", flush=True)
    sl = RaycastingScene(output_device="cuda"); sl.add_triangles(v, t); sl.commit()
    r1 = r[: 1 << 20].contiguous(); sl.list_intersections(r1); torch.cuda.synchronize()
    t0 = time.perf_counter(); ans = sl.list_intersections(r1); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{tag} canopy list_intersections 1M rays: {r1.shape[0]/dt/1e6:.0f} Mr/s wall ({ans['t_hit'].shape[0]} hits)", flush=True)
    del r, sl
if "c1" in want:
    v1, t1 = syn.qsm_tree_mesh(1)
    for sm, lm in ((8, 2), (1, 2), (8, 4), (1, 4), (4, 2), (16, 2)):
        c = RaycastingScene(output_device="cuda"); c.set_option("leaf_max", lm); c.set_option("split_max", sm); c.add_triangles(v1, t1)
        bms = min(c.commit() if k == 0 else (c.set_option("leaf_max", 1), c.set_option("leaf_max", lm), c.commit())[2] for k in range(3))
        st = c.stats()
        for G in (1000, 4000):
            r = grid_rays(c, syn.sun_direction(45, 135), G, G); o = outs(G * G); cnt = torch.empty(G * G, dtype=torch.int32, device="cuda")
            for var in (1, 2):
                c.set_option("traversal_variant", var)
                ms = gtime(lambda: _lib.check(L.qsmrt_cast_rays_2d(c._h, P(r), G, G, *[P(x) for x in o], None)), 5)
                mc = gtime(lambda: _lib.check(L.qsmrt_count_intersections(c._h, P(r), G * G, P(cnt), None)), 5)
                print(f"{tag} C1 split_max {sm} leaf_max {lm} (refs {st['num_references']}, build {bms:.3f} ms) grid {G} variant {var}: cast {G*G/ms/1e3:.0f} Mr/s ({ms:.3f} ms)  count {G*G/mc/1e3:.0f} Mr/s ({mc:.3f} ms)", flush=True)
            del r, o, cnt
