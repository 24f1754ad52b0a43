"""Which traversal kernel for which (scene, batch): cast_rays with the per-thread kernel (variant 1) and the persistent
kernel (variant 2) over cylinder-QSM plots and canopies of growing size and ray grids of growing size.  One line per
measurement (best of 5 launches, CUDA events); QSMRT_LIB selects the build (10 or 12 resident CTAs per SM).
    QSMRT_LIB=... python tools/probe_select.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
tag = os.path.basename(os.environ.get("QSMRT_LIB", "libqsmrt.so"))
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def gtime(f, reps=5):
    best = 1e30
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
def qsm_plot(k):                # k x k cylinder-QSM trees on a 6 m pitch
    vs, ts, base = [], [], 0
    for a in range(k * k):
        v, t = syn.qsm_tree_mesh(1 + a)
        v = v + np.asarray([6.0 * (a % k), 6.0 * (a // k), 0.0], np.float32)
        vs.append(v.astype(np.float32)); ts.append(t.astype(np.int64) + base); base += len(v)
    return np.concatenate(vs), np.concatenate(ts).astype(np.uint32)
scenes = [("qsm 1 tree", lambda: qsm_plot(1)), ("qsm 4 trees", lambda: qsm_plot(2)), ("qsm 16 trees", lambda: qsm_plot(4)),
          ("canopy 50k leaves", lambda: syn.canopy_mesh(2, 50_000)), ("canopy 250k leaves", lambda: syn.canopy_mesh(2, 250_000)),
          ("canopy 1M leaves", lambda: syn.canopy_mesh(2, 1_000_000))]
for name, make in scenes:
    v, t = make()
    s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit()
    st = s.stats()
    bvh_mb = (st["num_bvh_nodes"] * (32 if st["quantised_nodes"] else 64) + st["num_references"] * 48) / 1e6
    for G in (500, 1000, 2000, 4000):
        g = syn.parallel_ray_grid(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), syn.sun_direction(45, 135), G, G)
        r = torch.empty(G * G, 6, dtype=torch.float32, device="cuda")
        _lib.check(L.qsmrt_gen_parallel_rays(P(r), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
        o = [torch.empty(G * G, device="cuda"), torch.empty(G * G, dtype=torch.uint32, device="cuda"), torch.empty(G * G, dtype=torch.uint32, device="cuda"),
             torch.empty(G * G, 2, device="cuda"), torch.empty(G * G, 3, device="cuda")]
        res = {}
        for var in (1, 2):
            s.set_option("traversal_variant", var)
            f = lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), G, G, *[P(x) for x in o], None))
            f(); torch.cuda.synchronize()
            res[var] = gtime(f)
        print(f"{tag} {name}: refs {st['num_references']} bvh {bvh_mb:.0f} MB height {st['bvh_height']} rays {G*G}: v1 {res[1]:.3f} ms {G*G/res[1]/1e3:.0f} Mr/s | "
              f"v2 {res[2]:.3f} ms {G*G/res[2]/1e3:.0f} Mr/s | v1/v2 time {res[1]/res[2]:.3f}", flush=True)
        del r, o
    for n in (1 << 20, 1 << 22):        # incoherent rays through the scene's box
        r = torch.from_numpy(syn.random_rays(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), n, seed=7)).cuda()
        o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"),
             torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
        res = {}
        for var in (1, 2):
            s.set_option("traversal_variant", var)
            f = lambda: _lib.check(L.qsmrt_cast_rays(s._h, P(r), n, *[P(x) for x in o], None))
            f(); torch.cuda.synchronize()
            res[var] = gtime(f)
        print(f"{tag} {name}: random rays {n}: v1 {res[1]:.3f} ms {n/res[1]/1e3:.0f} Mr/s | v2 {res[2]:.3f} ms {n/res[2]/1e3:.0f} Mr/s | v1/v2 time {res[1]/res[2]:.3f}", flush=True)
        del r, o
    del s
