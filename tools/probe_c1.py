"""C1 (50k-triangle cylinder-QSM tree, 1M sun rays): one launch of each traversal kernel, for `ncu` (tail / issue
analysis of small batches).  Order of launches: cast v5, cast v1, count v5, cast v5 (16M rays), cast v1 (16M rays)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
v, t = syn.qsm_tree_mesh(1)
lm = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sm = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = RaycastingScene(output_device="cuda"); s.set_option("leaf_max", lm); s.set_option("split_max", sm); s.add_triangles(v, t); s.commit()
st = s.stats()
for G in (1000, 4000):
    g = syn.parallel_ray_grid(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), syn.sun_direction(45, 135), G, G)
    r = torch.empty(G * G, 6, dtype=torch.float32, device="cuda")
    _lib.check(L.qsmrt_gen_parallel_rays(P(r), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
    o = [torch.empty(G * G, device="cuda"), torch.empty(G * G, dtype=torch.uint32, device="cuda"), torch.empty(G * G, dtype=torch.uint32, device="cuda"),
         torch.empty(G * G, 2, device="cuda"), torch.empty(G * G, 3, device="cuda")]
    cnt = torch.empty(G * G, dtype=torch.int32, device="cuda")
    for var in (2, 1):
        s.set_option("traversal_variant", var)
        for rep in range(2):        # second launch of each = warm
            _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), G, G, *[P(x) for x in o], None)); torch.cuda.synchronize()
    s.set_option("traversal_variant", 2)
    _lib.check(L.qsmrt_count_intersections(s._h, P(r), G * G, P(cnt), None)); torch.cuda.synchronize()
    print("split_max", sm, "references", st["num_references"], "grid", G, "hit fraction", float(torch.isfinite(o[0]).float().mean()), "mean count", float(cnt.float().mean()), "max count", int(cnt.max()))
