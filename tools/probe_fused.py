"""Ad-hoc: scheduler thresholds of the fused kernels (sun exposure MODE 3, sky visibility MODE 4) on C2."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
v, t = syn.canopy_mesh(2, 1_000_000)
s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit()
lo, hi = v.min(0), v.max(0); G = 4000
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
cnt = torch.zeros(t.shape[0], dtype=torch.int32, device="cuda")
npts = 200_000
pts = torch.from_numpy(v[:npts].copy()).cuda(); free = torch.zeros(npts, dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
sweep = syn.hemisphere_sweep()
def t_sun():
    tot = 0.0
    for k in (3, 20, 41, 58):
        g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(*sweep[k]), G, G)
        best = 1e9
        for _ in range(2):
            e0.record(); _lib.check(L.qsmrt_sun_exposure(s._h, G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), P(cnt), None)); e1.record()
            torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        tot += best
    return 4 * G * G / tot / 1e3
def t_sky():
    best = 1e9
    for _ in range(2):
        e0.record(); _lib.check(L.qsmrt_sky_visibility(s._h, P(pts), None, npts, 5, C.c_float(1e-4), 0, 200, P(free), None)); e1.record()
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return npts * 200 / best / 1e3
for rf, wt, tm in ((12, 16, 1), (8, 16, 1), (6, 16, 1), (4, 16, 1), (12, 12, 1), (8, 12, 1), (4, 12, 1), (12, 20, 1), (8, 8, 1), (16, 16, 1), (12, 16, 4)):
    _lib.check(L.qsmrt_debug_set_tuning(rf, wt, tm, 0))
    print(f"refill {rf:2d} want {wt:2d} tri_min {tm}: sun {t_sun():6.0f} Mr/s   sky {t_sky():6.0f} Mr/s", flush=True)
_lib.check(L.qsmrt_debug_set_tuning(12, 16, 1, 0))
