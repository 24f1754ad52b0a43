"""Per-angle profile of the cast_rays kernel on C2 (the 64 solar angles of the sweep, 16M rays each), the data behind
bench.py's roofline: run once plainly with --counters (the kernel's own node / triangle fetch counts per angle ->
JSON), and once under `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,
dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_trace5 --csv` (one launch per angle, in sweep order);
`--merge counters.json launches.csv out.json` joins the two into profiles/r02_cast_rays_profile.json."""
import argparse, csv, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--counters", default=None, help="write the kernel's fetch counters per angle to this JSON")
ap.add_argument("--merge", nargs=3, default=None, metavar=("COUNTERS_JSON", "NCU_CSV", "OUT_JSON"))
ap.add_argument("--grid", type=int, default=4000)
a = ap.parse_args()
if a.merge:
    cnt = json.load(open(a.merge[0]))
    rows = list(csv.reader(open(a.merge[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in data:
        if len(r) > vi and "k_trace5<0" in r[ki]:
            per.setdefault(int(r[ii]), {})[r[mi]] = float(r[vi].replace(",", ""))
    launches = [per[k] for k in sorted(per)][1:]          # launch 0 is the warm-up
    assert len(launches) == len(cnt["angles"]), (len(launches), len(cnt["angles"]))
    n = cnt["rays_per_angle"]
    out = {"source": "tools/profile_angles.py on a B200: one k_trace5<0,0,1> launch per solar angle of the C2 sweep (%d rays each); ncu --metrics "
                     "gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none; fetch counters from a plain run with set_option('counters', 1)" % n,
           "rays_per_angle": n, "angles": []}
    for ang, m in zip(cnt["angles"], launches):
        out["angles"].append({"elevation": ang["elevation"], "azimuth": ang["azimuth"], "nodes_per_ray": ang["nodes_per_ray"], "tris_per_ray": ang["tris_per_ray"],
                              "kernel_ms": ang["kernel_ms"], "ncu_ms": m["gpu__time_duration.sum"] / 1e6, "warp_inst": m["smsp__inst_executed.sum"],
                              "thread_inst": m["smsp__thread_inst_executed.sum"], "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]})
    A = out["angles"]
    out["warp_inst_per_ray"] = sum(x["warp_inst"] for x in A) / (n * len(A))
    out["lanes_per_inst"] = sum(x["thread_inst"] for x in A) / sum(x["warp_inst"] for x in A)
    out["fetch_bytes_per_ray"] = sum(32 * x["nodes_per_ray"] + 48 * x["tris_per_ray"] for x in A) / len(A)
    out["dram_bytes_per_launch"] = sum(x["dram_bytes"] for x in A) / len(A)
    json.dump(out, open(a.merge[2], "w"), indent=1)
    print("wrote", a.merge[2], "warp inst/ray %.1f lanes %.2f fetch B/ray %.0f dram MB/launch %.0f" % (
        out["warp_inst_per_ray"], out["lanes_per_inst"], out["fetch_bytes_per_ray"], out["dram_bytes_per_launch"] / 1e6))
    sys.exit(0)
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load(); G = a.grid; n = G * G
v, t = syn.canopy_mesh(2, 1_000_000)
s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit()
st = s.stats(); lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
rays = torch.empty(n, 6, dtype=torch.float32, device="cuda")
o = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
def cast():
    _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), G, G, *[P(x) for x in o], None))
res = []
sweep = syn.hemisphere_sweep()
for k, (el, az) in enumerate([sweep[0]] + sweep):          # the first launch is a warm-up
    g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), G, G)
    _lib.check(L.qsmrt_gen_parallel_rays(P(rays), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
    torch.cuda.synchronize()
    e0.record(); cast(); e1.record(); torch.cuda.synchronize()
    if k == 0: continue
    row = {"elevation": el, "azimuth": az, "kernel_ms": e0.elapsed_time(e1)}
    if a.counters:
        s.set_option("counters", 1); cast(); c = s.counters(); s.set_option("counters", 0)
        row["nodes_per_ray"], row["tris_per_ray"] = c[0] / n, c[1] / n
    res.append(row)
if a.counters:
    json.dump({"rays_per_angle": n, "angles": res}, open(a.counters, "w"), indent=1)
print("mean kernel ms %.3f -> %.0f Mrays/s" % (np.mean([r["kernel_ms"] for r in res]), n / np.mean([r["kernel_ms"] for r in res]) / 1e3))
