"""Ad-hoc: throughput of the fused environmental drivers on the C2 canopy (not product code)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, environment as env
v, t = syn.canopy_mesh(2, 1_000_000)
s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); print("build ms", s.commit())
def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, time.perf_counter() - t0
sweep = syn.hemisphere_sweep()
env.sun_exposure(s, sweep[:2], grid=(4000, 4000))
r, dt = timed(lambda: env.sun_exposure(s, sweep, grid=(4000, 4000)))
print(f"sun sweep 64 angles x 16M rays: {dt*1e3:.1f} ms  {r['rays']/dt/1e6:.0f} Mrays/s  sunlit-ray total {int(r['counts'].sum())}")
# C5: sky MC from leaf vertices (every 4th vertex = one per leaf), normals = leaf normals
tri = t.reshape(-1, 2, 3)[:, 0]
p0, p1, p2 = v[tri[:, 0]], v[tri[:, 1]], v[tri[:, 2]]
n = np.cross(p1 - p0, p2 - p0); n /= np.linalg.norm(n, axis=1, keepdims=True)
pts = torch.from_numpy(p0).cuda(); nrm = torch.from_numpy(n.astype(np.float32)).cuda()
env.sky_gap_fraction(s, pts[:1000], nrm[:1000], n_dirs=10)
for nd in (100,):
    gap, dt = timed(lambda: env.sky_gap_fraction(s, pts, nrm, n_dirs=nd, seed=5))
    print(f"sky MC {pts.shape[0]} points x {nd} dirs = {pts.shape[0]*nd/1e6:.0f}M rays: {dt*1e3:.1f} ms  {pts.shape[0]*nd/dt/1e6:.0f} Mrays/s  mean gap {gap.mean().item():.4f}")
r, dt = timed(lambda: env.rain_interception(s, 20.0, 0.0, grid=(4000, 4000)))
print(f"rain 16M rays count_intersections: {dt*1e3:.1f} ms {r['rays']/dt/1e6:.0f} Mrays/s intercepted {r['intercepted_fraction']:.3f} mean layers {r['mean_layers']:.2f}")
import ctypes as C
from pyqsm_b200 import _lib
L = _lib.load()
st = s.stats(); lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x]); P = lambda x: C.c_void_p(x.data_ptr())
n = 4000 * 2000
buf = torch.empty(n, 6, device="cuda"); cnt = torch.empty(n, dtype=torch.int32, device="cuda"); occ = torch.empty(n, dtype=torch.uint8, device="cuda")
for el in (70.0, 45.0, 20.0):
    g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, 0.0), 4000, 4000)
    _, t0 = timed(lambda: _lib.check(L.qsmrt_gen_parallel_rays(P(buf), 4000, 2000, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None)))
    _, t1 = timed(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(buf), n, P(cnt), None)))
    _, t1b = timed(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(buf), n, P(cnt), None)))
    _, t2 = timed(lambda: torch.bincount(cnt.clamp(max=255).to(torch.int64), minlength=256))
    _, t3 = timed(lambda: _lib.check(L.qsmrt_test_occlusions(s._h, P(buf), n, 0.0, float("inf"), P(occ), None)))
    print(f"el {el}: gen {t0*1e3:.2f} ms count {t1*1e3:.2f}/{t1b*1e3:.2f} ms ({n/t1b/1e6:.0f} Mr/s) bincount {t2*1e3:.2f} ms occl {t3*1e3:.2f} ms max {cnt.max().item()} mean {cnt.float().mean().item():.2f} n>32: {(cnt>32).sum().item()}")
