"""A/B probe on the big configs for one build of libqsmrt (QSMRT_LIB): C3 (10M triangles, count_intersections, 10M slanted
rays) and C4 (50M triangles, 64-byte nodes, cast_rays 16M rays).    python tools/probe_big.py [c3 c4]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load(); tag = os.path.basename(os.environ.get("QSMRT_LIB", "libqsmrt.so")); want = sys.argv[1:] or ["c3", "c4"]
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def gtime(f, reps=3):
    best = 1e30
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
def scene(v, t):
    s = RaycastingScene(output_device="cuda"); s.add_triangles(torch.from_numpy(v).cuda(), torch.from_numpy(t.view(np.int32)).cuda().view(torch.uint32)); s.commit(); return s
def grid(s, d, nu, nv):
    st = s.stats(); g = syn.parallel_ray_grid(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), d, nu, nv)
    r = torch.empty(nu * nv, 6, dtype=torch.float32, device="cuda")
    _lib.check(L.qsmrt_gen_parallel_rays(P(r), nu, nv, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None)); return r
if "c3" in want:
    v, t = syn.plot_mesh(3, 5, 1_000_000, 14.0); s = scene(v, t)
    r = grid(s, syn.sun_direction(70.0, 0.0), 10000, 10000)[40_000_000:50_000_000].contiguous(); cnt = torch.empty(r.shape[0], dtype=torch.int32, device="cuda")
    ms = gtime(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(r), r.shape[0], P(cnt), None)))
    o = [torch.empty(r.shape[0], device="cuda"), torch.empty(r.shape[0], dtype=torch.uint32, device="cuda")]
    mc = gtime(lambda: _lib.check(L.qsmrt_cast_rays(s._h, P(r), r.shape[0], P(o[0]), None, P(o[1]), None, None, None)))
    print(f"{tag} C3 10M tris, 10M rays (rows 4000-5000): count {r.shape[0]/ms/1e3:.0f} Mr/s ({ms:.2f} ms), cast (t, prim) {r.shape[0]/mc/1e3:.0f} Mr/s", flush=True)
    del s, r, cnt, o
if "c4" in want:
    v, t = syn.plot_mesh(4, 25, 1_000_000, 40.0); s = scene(v, t); st = s.stats()
    r = grid(s, syn.sun_direction(60, 30), 4000, 4000); n = r.shape[0]
    o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
    for q in (0, 1):
        s.set_option("quant_threshold", 100.0 if q else 0.15); s.commit()
        ms = gtime(lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), 4000, 4000, *[P(x) for x in o], None)))
        print(f"{tag} C4 50M tris, 16M rays, quantised nodes {s.stats()['quantised_nodes']}: cast {n/ms/1e3:.0f} Mr/s ({ms:.2f} ms), build {s.stats()['build_ms']:.2f} ms", flush=True)
