"""What does dealing a step's N angles in row blocks to the ranks cost one rank?  Emulates rank 0 of N = 8 on ONE GPU:
(a) eight single-angle launches, (b) the composite 16M-ray buffer of bench.py for several row-block sizes, each with
tile rows centre-out and in memory order; also the canopy count kernel for several hit-set sizes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load(); G = 4000; n = G * G
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def gtime(f, reps=3):
    best = 1e30
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
v, t = syn.canopy_mesh(2, 1_000_000)
s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit(); st = s.stats()
lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
sweep = syn.hemisphere_sweep(); angles = [sweep[(3 * k) % 64] for k in range(8)]
o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
full = [torch.empty(n, 6, dtype=torch.float32, device="cuda") for _ in range(8)]
for k, (el, az) in enumerate(angles):
    g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), G, G)
    _lib.check(L.qsmrt_gen_parallel_rays(P(full[k]), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
cast = lambda r: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), G, G, *[P(x) for x in o], None))
for order in (1, 0):
    s.set_option("tile_order", order)
    ts = [gtime(lambda: cast(full[k])) for k in range(8)]
    print(f"tile_order {order}: single angles mean {np.mean(ts):.3f} ms ({' '.join('%.2f' % x for x in ts)})", flush=True)
    buf = torch.empty(n, 6, dtype=torch.float32, device="cuda")
    for rb in (4, 20, 100, 500):
        nb = G // rb // 8
        for k in range(8):
            buf.view(8, nb, rb * G * 6)[k].copy_(full[k].view(nb, 8, rb * G * 6)[:, 0])
        print(f"tile_order {order}: composite of 8 angles, row block {rb}: {gtime(lambda: cast(buf)):.3f} ms", flush=True)
    del buf
r = full[0][: 4000 * 1000].contiguous(); cnt = torch.empty(r.shape[0], dtype=torch.int32, device="cuda")
for cs in (32, 24, 16, 12, 8):
    s.set_option("count_set", cs)
    ms = gtime(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(r), r.shape[0], P(cnt), None)))
    print(f"canopy count 4M rays, hit set {cs}: {r.shape[0]/ms/1e3:.0f} Mr/s (max count {int(cnt.max())})", flush=True)
