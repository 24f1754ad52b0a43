"""Resident CTAs per SM for small batches (QSMRT_OPT_CTAS_PER_SM): C1 cast_rays at 1M rays with 4..10 CTAs per SM.  Not product code."""
import ctypes as C, os, sys
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/pyqsm_b200') else os.getcwd())
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def gtime(f, reps=5):
    best = 1e30
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
v, t = syn.qsm_tree_mesh(1)
for lm in (2, 4):
    s = RaycastingScene(output_device="cuda"); s.set_option("leaf_max", lm); s.add_triangles(v, t); s.commit(); st = s.stats()
    for G in (1000, 2000):
        g = syn.parallel_ray_grid(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), syn.sun_direction(45, 135), G, G)
        r = torch.empty(G * G, 6, dtype=torch.float32, device="cuda")
        _lib.check(L.qsmrt_gen_parallel_rays(P(r), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
        o = [torch.empty(G * G, device="cuda"), torch.empty(G * G, dtype=torch.uint32, device="cuda"), torch.empty(G * G, dtype=torch.uint32, device="cuda"), torch.empty(G * G, 2, device="cuda"), torch.empty(G * G, 3, device="cuda")]
        cnt = torch.empty(G * G, dtype=torch.int32, device="cuda")
        for ctas in (0, 8, 6, 5, 4, 3, 2):
            s.set_option("ctas_per_sm", ctas)
            ms = gtime(lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), G, G, *[P(x) for x in o], None)))
            mc = gtime(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(r), G * G, P(cnt), None)))
            print(f"C1 leaf_max {lm} grid {G} ctas/SM {ctas}: cast {ms:.3f} ms {G*G/ms/1e3:.0f} Mr/s | count {mc:.3f} ms {G*G/mc/1e3:.0f} Mr/s", flush=True)
