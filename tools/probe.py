"""Ad-hoc GPU probe (not part of the product): times build + queries on a config,
A/B over traversal variants (1|2) and ray->thread mappings (linear|2d)."""
import argparse, ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--leaves", type=int, default=1_000_000)
ap.add_argument("--grid", type=int, default=4000)
ap.add_argument("--angles", type=int, default=3)
ap.add_argument("--mesh", default="canopy")
ap.add_argument("--others", action="store_true")
a = ap.parse_args()
v, t = syn.canopy_mesh(2, a.leaves) if a.mesh == "canopy" else syn.qsm_tree_mesh(1)
s = RaycastingScene(output_device="cuda")
s.add_triangles(v, t)
for i in range(2):
    s2 = RaycastingScene(output_device="cuda"); s2.add_triangles(v, t); ms = s2.commit(); st = s2.stats()
    print("build_ms", round(ms, 3), "sort_ms", round(st["sort_ms"], 3), "nodes", st["num_bvh_nodes"], "leaves", st["num_bvh_leaves"]); del s2
s.commit()
L = _lib.load()
n = a.grid * a.grid
rays = torch.empty(n, 6, dtype=torch.float32, device="cuda")
o = dict(t=torch.empty(n, device="cuda"), g=torch.empty(n, dtype=torch.uint32, device="cuda"), p=torch.empty(n, dtype=torch.uint32, device="cuda"),
         uv=torch.empty(n, 2, device="cuda"), nr=torch.empty(n, 3, device="cuda"))
lo, hi = v.min(0), v.max(0)
sweep = syn.hemisphere_sweep()
P = lambda x: C.c_void_p(x.data_ptr())
F3 = lambda x: (C.c_float * 3)(*x)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def timeit(f, reps=3):
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for k in range(a.angles):
    el, az = sweep[(k * 9) % 64]
    g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), a.grid, a.grid)
    _lib.check(L.qsmrt_gen_parallel_rays(P(rays), a.grid, a.grid, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), st))
    res = []
    ref = None
    for var in (1, 2):
        _lib.check(L.qsmrt_debug_set_variant(var))
        for mode in ("lin", "2d"):
            if mode == "lin":
                f = lambda: _lib.check(L.qsmrt_cast_rays(s._h, P(rays), n, P(o["t"]), P(o["g"]), P(o["p"]), P(o["uv"]), P(o["nr"]), st))
            else:
                f = lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), a.grid, a.grid, P(o["t"]), P(o["g"]), P(o["p"]), P(o["uv"]), P(o["nr"]), st))
            ms = timeit(f)
            cur = (o["t"].clone(), o["p"].clone(), o["uv"].clone())
            if ref is None: ref = cur
            same = all(torch.equal(x, y) for x, y in zip(ref, cur))
            res.append(f"v{var}/{mode} {ms:.2f} ms {n/ms/1e3:.0f} Mr/s {'=' if same else 'DIFF'}")
    print(f"el {el:.0f} az {az:.0f} hit {torch.isfinite(o['t']).float().mean().item():.3f} | " + " | ".join(res), flush=True)
    if a.others:
        cnt = torch.empty(n, dtype=torch.int32, device="cuda"); occ = torch.empty(n, dtype=torch.uint8, device="cuda")
        cms = timeit(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(rays), n, P(cnt), st)), 2)
        oms = timeit(lambda: _lib.check(L.qsmrt_test_occlusions(s._h, P(rays), n, 0.0, float("inf"), P(occ), st)), 2)
        print(f"   count {cms:.2f} ms {n/cms/1e3:.0f} Mr/s max {cnt.max().item()} | occl {oms:.2f} ms {n/oms/1e3:.0f} Mr/s", flush=True)
