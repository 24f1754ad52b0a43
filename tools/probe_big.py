"""Ad-hoc: the large BASELINE configs (C3 10M triangles, C4 50M triangles) build and trace (not product code)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
for ncan, pitch in ((5, 14.0), (25, 40.0)):
    t0 = time.time(); v, t = syn.plot_mesh(3 if ncan == 5 else 4, ncan, 1_000_000, pitch); print(f"mesh {t.shape[0]/1e6:.0f}M tris host gen {time.time()-t0:.1f}s", flush=True)
    vd, td = torch.from_numpy(v).cuda(), torch.from_numpy(t.view(np.int32)).cuda().view(torch.uint32)
    builds = []
    for i in range(3):
        s = RaycastingScene(output_device="cuda"); s.add_triangles(vd, td); builds.append(s.commit())
        if i < 2: del s
    st = s.stats(); print("build ms", [round(b, 2) for b in builds], "sort", round(st["sort_ms"], 2), "height", st["bvh_height"], "quant", st["quantised_nodes"], "bvh MB", st["bvh_bytes"] / 1e6, "mem GB", torch.cuda.memory_allocated() / 1e9, flush=True)
    lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    n = 4000 * 4000
    rays = torch.empty(n, 6, device="cuda"); o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for el, az in ((70.0, 0.0), (45.0, 135.0)):
        g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), 4000, 4000)
        _lib.check(L.qsmrt_gen_parallel_rays(P(rays), 4000, 4000, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
        for rep in range(2):
            e0.record(); _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), 4000, 4000, *[P(x) for x in o], None)); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        e0.record(); _lib.check(L.qsmrt_count_intersections(s._h, P(rays), n, P(cnt), None)); e1.record(); torch.cuda.synchronize()
        cms = e0.elapsed_time(e1)
        print(f"  el {el:.0f}: cast {ms:.2f} ms {n/ms/1e3:.0f} Mr/s hit {torch.isfinite(o[0]).float().mean().item():.3f} | count {cms:.2f} ms {n/cms/1e3:.0f} Mr/s max {cnt.max().item()} mean {cnt.float().mean().item():.2f}", flush=True)
    del s, vd, td, rays, o, cnt; torch.cuda.empty_cache()
