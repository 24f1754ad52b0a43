"""Ad-hoc: 32-byte quantised nodes vs 64-byte fp32 nodes on a wide plot (200 m: the 16-bit grid is 3 mm), where the
default threshold rejects them.  cast_rays, 16M sun rays, several angles."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
ncan = int(sys.argv[1]) if len(sys.argv) > 1 else 9
v, t = syn.plot_mesh(4, ncan, 1_000_000, 40.0)
vd, td = torch.from_numpy(v).cuda(), torch.from_numpy(t.view(np.int32)).cuda().view(torch.uint32)
lo, hi = v.min(0), v.max(0)
G = 4000; n = G * G
rays = torch.empty(n, 6, dtype=torch.float32, device="cuda")
o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"),
     torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*x)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
ref = {}
for frac in (0.15, 0.5, 1.0):
    _lib.check(L.qsmrt_debug_set_quant_threshold(C.c_float(frac)))
    s = RaycastingScene(output_device="cuda"); s.add_triangles(vd, td); ms = s.commit(); st = s.stats()
    line = f"frac {frac:4.2f} quantised {st['quantised_nodes']} build {ms:.2f} ms |"
    for (el, az) in ((30.0, 45.0), (60.0, 135.0), (80.0, 0.0)):
        g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), G, G)
        _lib.check(L.qsmrt_gen_parallel_rays(P(rays), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
        best = 1e9
        for _ in range(3):
            e0.record(); _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), G, G, *[P(x) for x in o], None)); e1.record()
            torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        key = (el, az); cur = (o[0].clone(), o[2].clone())
        same = key not in ref or (torch.equal(ref[key][0], cur[0]) and torch.equal(ref[key][1], cur[1]))
        ref.setdefault(key, cur)
        line += f" el{el:.0f}: {best:.2f} ms {n / best / 1e3:.0f} Mr/s {'=' if same else 'DIFF'} |"
    print(line, flush=True)
    del s
_lib.check(L.qsmrt_debug_set_quant_threshold(C.c_float(0.0)))
