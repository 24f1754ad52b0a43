"""Ad-hoc GPU probe (not part of the product): times build + cast_rays on a config."""
import argparse, ctypes as C, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--leaves", type=int, default=1_000_000)
ap.add_argument("--grid", type=int, default=4000)
ap.add_argument("--angles", type=int, default=4)
ap.add_argument("--mesh", default="canopy")
a = ap.parse_args()
t0 = time.time()
if a.mesh == "canopy":
    v, t = syn.canopy_mesh(2, a.leaves)
else:
    v, t = syn.qsm_tree_mesh(1)
print("mesh", v.shape, t.shape, f"{time.time()-t0:.1f}s", flush=True)
s = RaycastingScene(output_device="cuda")
s.add_triangles(v, t)
for i in range(3):
    s2 = RaycastingScene(output_device="cuda"); s2.add_triangles(v, t); ms = s2.commit(); st = s2.stats()
    print("build_ms", ms, "sort_ms", st["sort_ms"], "nodes", st["num_bvh_nodes"], "leaves", st["num_bvh_leaves"]); del s2
s.commit()
L = _lib.load()
n = a.grid * a.grid
rays = torch.empty(n, 6, dtype=torch.float32, device="cuda")
outs = dict(t=torch.empty(n, device="cuda"), g=torch.empty(n, dtype=torch.uint32, device="cuda"), p=torch.empty(n, dtype=torch.uint32, device="cuda"),
            uv=torch.empty(n, 2, device="cuda"), nr=torch.empty(n, 3, device="cuda"))
lo, hi = v.min(0), v.max(0)
sweep = syn.hemisphere_sweep()
P = lambda x: C.c_void_p(x.data_ptr())
F3 = lambda x: (C.c_float * 3)(*x)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for k in range(a.angles):
    el, az = sweep[(k * 9) % 64]
    o0, du, dv, d = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), a.grid, a.grid)
    _lib.check(L.qsmrt_gen_parallel_rays(P(rays), a.grid, a.grid, F3(o0), F3(du), F3(dv), F3(d), st))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    for rep in range(2):
        e0.record()
        _lib.check(L.qsmrt_cast_rays(s._h, P(rays), n, P(outs["t"]), P(outs["g"]), P(outs["p"]), P(outs["uv"]), P(outs["nr"]), st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    hit = torch.isfinite(outs["t"]).float().mean().item()
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    e0.record(); _lib.check(L.qsmrt_count_intersections(s._h, P(rays), n, P(cnt), st)); e1.record(); torch.cuda.synchronize()
    cms = e0.elapsed_time(e1)
    occ = torch.empty(n, dtype=torch.uint8, device="cuda")
    e0.record(); _lib.check(L.qsmrt_test_occlusions(s._h, P(rays), n, 0.0, float("inf"), P(occ), st)); e1.record(); torch.cuda.synchronize()
    oms = e0.elapsed_time(e1)
    print(f"el {el:.0f} az {az:.0f}: cast {ms:.2f} ms = {n/ms/1e3:.1f} Mrays/s hit {hit:.3f} | count {cms:.2f} ms {n/cms/1e3:.1f} Mrays/s max {cnt.max().item()} | occl {oms:.2f} ms {n/oms/1e3:.1f} Mrays/s", flush=True)
