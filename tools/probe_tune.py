"""Ad-hoc: sweep the tuning of the persistent kernels on C2 (not product code)."""
import argparse, ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
ap = argparse.ArgumentParser(); ap.add_argument("--grid", type=int, default=4000); ap.add_argument("--angles", type=int, default=2)
ap.add_argument("--mesh", default="c2", choices=["c1", "c2"])
ap.add_argument("--combos", default="2:12,12,1,2,0,1;2:12,12,1,2,0,0;2:12,12,1,4,0,1;2:8,8,1,2,0,1;2:16,16,1,2,0,1")
a = ap.parse_args()
v, t = syn.canopy_mesh(2, 1_000_000) if a.mesh == "c2" else syn.qsm_tree_mesh(1)
s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit(); print("stats", s.stats())
L = _lib.load(); n = a.grid * a.grid
rays = torch.empty(n, 6, dtype=torch.float32, device="cuda")
o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*x)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
lo, hi = v.min(0), v.max(0); sweep = syn.hemisphere_sweep()
def run():
    _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), a.grid, a.grid, *[P(x) for x in o], st))
def timeit(reps=3):
    best = 1e9
    for _ in range(reps):
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
cur_lm = 2
for k in range(a.angles):
    el, az = sweep[(k * 27 + 5) % 64] if a.mesh == "c2" else (45.0, 135.0 + 30.0 * k)
    g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), a.grid, a.grid)
    _lib.check(L.qsmrt_gen_parallel_rays(P(rays), a.grid, a.grid, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), st))
    s.set_option("traversal_variant", 1); ms3 = timeit(); ref = (o[0].clone(), o[2].clone(), o[3].clone())
    print(f"el {el:.0f} az {az:.0f}: v1 {ms3:.2f} ms {n/ms3/1e3:.0f} Mr/s")
    for combo in a.combos.split(";"):
        var, rest = combo.split(":"); rf, wt, tm, lm, *npth = map(int, rest.split(","))
        s.set_option("node_path", npth[0] if npth else 0)
        s.set_option("quantised_nodes", npth[1] if len(npth) > 1 else 1)
        if lm != cur_lm:
            s.set_option("leaf_max", lm); cur_lm = lm; s.commit()       # a changed builder option rebuilds
        s.set_option("traversal_variant", int(var))
        for k, x in (("refill", rf), ("want", wt), ("tri_min", tm), ("counters", 0)): s.set_option(k, x)
        ms = timeit()
        same = torch.equal(ref[0], o[0]) and torch.equal(ref[1], o[2]) and torch.equal(ref[2], o[3])
        s.set_option("counters", 1); run(); cs = s.counters(); s.set_option("counters", 0)
        class _V:  # noqa: E701
            def __init__(self, v): self.value = v
        nn, nt = _V(cs[0]), _V(cs[1])
        if cs[2]: print(f"      node-phase iters/ray {cs[2]*32/n:.1f} lanes step {cs[3]/cs[2]:.1f} idle {cs[4]/cs[2]:.1f} leaf2 {cs[5]/cs[2]:.1f} done {cs[6]/cs[2]:.1f} | tri-phase iters/ray {cs[7]*32/n:.1f} lanes {cs[8]/max(cs[7],1):.1f}")
        print(f"   v{var} refill {rf:2d} want {wt:2d} trimin {tm:2d} leafmax {lm} path {npth[0] if npth else 0} quant {npth[1] if len(npth) > 1 else 1}: {ms:.2f} ms {n/ms/1e3:.0f} Mr/s {'=' if same else 'DIFF'}  nodes/ray {nn.value/n:.1f} tris/ray {nt.value/n:.2f}", flush=True)
