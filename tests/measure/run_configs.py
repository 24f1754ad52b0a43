"""Runs the five BASELINE.json configurations at full size on one B200 and prints the results table of
BASELINE.md section 7 (Mrays/s, canonical bytes per ray, roofline, build ms, CPU oracle Mrays/s, parity).
Test / measurement infrastructure: uses the oracle as checker and CPU baseline.

    python tests/measure/run_configs.py [c1 c2 c3 c4 c5] > profiles/r02_configs.json

Run it under `ncu --metrics ... -k regex:k_trace5` as well to see which resource each configuration's launches
load (issue slots, ALU pipe, L2, DRAM): profiles/r02_configs_ncu.txt.
"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import oracle
from pyqsm_b200 import RaycastingScene, synthetic as syn, environment as env, _lib

L = _lib.load()
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")) else 6552.0
want = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c4", "c5"]
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
rows = []

def gpu_time(f, reps=3):
    best = 1e30
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

def gen_grid(scene, direction, nu, nv, rows=None, row0=0):
    st = scene.stats(); lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
    g = syn.parallel_ray_grid(lo, hi, direction, nu, nv)
    nr = nv if rows is None else rows
    o0 = (g[0].astype(np.float64) + row0 * g[2].astype(np.float64)).astype(np.float32)
    rays = torch.empty(nr * nu, 6, dtype=torch.float32, device="cuda")
    _lib.check(L.qsmrt_gen_parallel_rays(P(rays), nu, nr, F3(o0), F3(g[1]), F3(g[2]), F3(g[3]), None))
    return rays

def report(**kw):
    rows.append(kw); print(json.dumps(kw), flush=True)

def build(v, t):
    vd, td = torch.from_numpy(v).cuda(), torch.from_numpy(t.view(np.int32)).cuda().view(torch.uint32)
    bs = []
    for i in range(3):
        s = RaycastingScene(output_device="cuda"); s.add_triangles(vd, td); bs.append(s.commit())
        if i < 2: del s
    return s, min(bs)

def oracle_scene(v, t):
    o = oracle.OracleScene(); o.add_triangles(v, t); t0 = time.perf_counter(); o.commit(); return o, (time.perf_counter() - t0) * 1e3

if "c1" in want:
    v, t = syn.qsm_tree_mesh(1); s, bms = build(v, t); o, obms = oracle_scene(v, t)
    rays = gen_grid(s, syn.sun_direction(45, 135), 1000, 1000); n = rays.shape[0]
    out = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = gpu_time(lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), 1000, 1000, *[P(x) for x in out], None)), 5)
    cms = gpu_time(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(rays), n, P(cnt), None)), 5)
    rh = rays.cpu().numpy(); t0 = time.perf_counter(); ref = o.cast_rays(rh, 1); cpu_s = time.perf_counter() - t0; nn, nt = o.last_counters
    t0 = time.perf_counter(); refc = o.count_intersections(rh, 1); cpu_c = time.perf_counter() - t0; cn, ct = o.last_counters
    same = bool(np.array_equal(out[2].cpu().numpy(), ref["primitive_ids"]) and np.array_equal(out[0].cpu().numpy(), ref["t_hit"]) and np.array_equal(cnt.cpu().numpy(), refc))
    b = 24 + 32 + 32 * nn / n + 48 * nt / n; bc = 24 + 4 + 32 * cn / n + 48 * ct / n
    report(config="C1 cast_rays", triangles=int(t.shape[0]), rays=n, mrays_s=n / ms / 1e3, b_ray=b, roofline_mrays_s=HBM * 1e3 / b, frac_canonical_hbm=(n / ms / 1e3) / (HBM * 1e3 / b), build_ms=bms, cpu_mrays_s=n / cpu_s / 1e6, cpu_threads=o.num_threads, cpu_build_ms=obms, parity="bit-identical (all 1M rays)" if same else "MISMATCH")
    report(config="C1 count_intersections", triangles=int(t.shape[0]), rays=n, mrays_s=n / cms / 1e3, b_ray=bc, roofline_mrays_s=HBM * 1e3 / bc, frac_canonical_hbm=(n / cms / 1e3) / (HBM * 1e3 / bc), cpu_mrays_s=n / cpu_c / 1e6, parity="bit-identical (all 1M rays)" if same else "MISMATCH")
    del s, o

if "c2" in want or "c5" in want:
    v2, t2 = syn.canopy_mesh(2, 1_000_000); s2, bms2 = build(v2, t2); o2, obms2 = oracle_scene(v2, t2)
if "c2" in want:
    sweep = syn.hemisphere_sweep(); env.sun_exposure(s2, sweep[:2], grid=(4000, 4000)); torch.cuda.synchronize()
    t0 = time.perf_counter(); r = env.sun_exposure(s2, sweep, grid=(4000, 4000)); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    cc = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "baseline", "canonical_counters.json")))["summary"]["c2_canopy_2m_cast"]
    rays = gen_grid(s2, syn.sun_direction(40, 135), 4000, 4000); n = rays.shape[0]
    out = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
    ms = gpu_time(lambda: _lib.check(L.qsmrt_cast_rays_2d(s2._h, P(rays), 4000, 4000, *[P(x) for x in out], None)))
    sub = rays[::16].cpu().numpy(); t0 = time.perf_counter(); ref = o2.cast_rays(sub, 1); cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(out[2][::16].cpu().numpy(), ref["primitive_ids"]) and np.array_equal(out[0][::16].cpu().numpy(), ref["t_hit"]))
    report(config="C2 cast_rays (one angle, API)", triangles=int(t2.shape[0]), rays=n, mrays_s=n / ms / 1e3, b_ray=cc["b_ray"], roofline_mrays_s=HBM * 1e3 / cc["b_ray"], frac_canonical_hbm=(n / ms / 1e3) / (HBM * 1e3 / cc["b_ray"]), build_ms=bms2, cpu_mrays_s=len(sub) / cpu_s / 1e6, cpu_threads=o2.num_threads, cpu_build_ms=obms2, parity="bit-identical (1M-ray subsample)" if same else "MISMATCH")
    report(config="C2 fused sun sweep, 64 angles x 16M rays (one launch, wall clock)", rays=r["rays"], mrays_s=r["rays"] / dt / 1e6, seconds=dt, sunlit_rays=int(r["counts"].sum()))
    del rays, out

if "c5" in want:
    tri = t2.reshape(-1, 2, 3)[:, 0]; p0, p1, p2 = v2[tri[:, 0]], v2[tri[:, 1]], v2[tri[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p0); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts, nd = torch.from_numpy(p0).cuda(), torch.from_numpy(nrm.astype(np.float32)).cuda()
    env.sky_gap_fraction(s2, pts, nd, n_dirs=8); torch.cuda.synchronize()          # warm-up at full size (allocates the point-order scratch)
    t0 = time.perf_counter(); gap = env.sky_gap_fraction(s2, pts, nd, n_dirs=1000, seed=5); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    base = 517_000                                   # a block of 1000 points OF THE FULL LAUNCH: same place in the sample via point_base
    rays = env.hemisphere_rays(pts[base:base + 1000], nd[base:base + 1000], n_dirs=1000, seed=5, point_base=base)
    rh = rays.cpu().numpy(); t0 = time.perf_counter(); occ = o2.test_occlusions(rh, mode=1); cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(np.rint(gap[base:base + 1000].cpu().numpy() * 1000).astype(np.int64), 1000 - occ.reshape(1000, 1000).sum(1)))
    report(config="C5 sky Monte-Carlo gap fraction (1M points x 1000 directions, rays never materialised)", rays=int(pts.shape[0]) * 1000, mrays_s=pts.shape[0] * 1000 / dt / 1e6, seconds=dt, mean_gap=float(gap.mean()), cpu_mrays_s=len(rh) / cpu_s / 1e6, parity="identical to oracle occlusion on a 1000-point block of the same launch (1M rays)" if same else "MISMATCH")
if "c2" in want or "c5" in want:
    del s2, o2

if "c3" in want:
    v, t = syn.plot_mesh(3, 5, 1_000_000, 14.0); s, bms = build(v, t); o, obms = oracle_scene(v, t)
    d = syn.sun_direction(70.0, 0.0); nu = nv = 10000; chunk = 1000; total_ms = 0.0; hist = torch.zeros(64, dtype=torch.int64, device="cuda")
    cnt = torch.empty(nu * chunk, dtype=torch.int32, device="cuda"); sub_rays = []; sub_cnt = []
    for r0 in range(0, nv, chunk):
        rays = gen_grid(s, d, nu, nv, rows=chunk, row0=r0)
        total_ms += gpu_time(lambda: _lib.check(L.qsmrt_count_intersections(s._h, P(rays), nu * chunk, P(cnt), None)), 1)
        hist += torch.bincount(cnt.clamp(max=63).to(torch.int64), minlength=64)
        sub_rays.append(rays[::100].cpu().numpy()); sub_cnt.append(cnt[::100].cpu().numpy())
    sub = np.concatenate(sub_rays); subc = np.concatenate(sub_cnt)
    t0 = time.perf_counter(); refc = o.count_intersections(sub, 1); cpu_s = time.perf_counter() - t0; cn, ct = o.last_counters
    bc = 24 + 4 + 32 * cn / len(sub) + 48 * ct / len(sub); n = nu * nv
    report(config="C3 rain count_intersections (10M triangles, 100M rays 20 deg off vertical)", triangles=int(t.shape[0]), rays=n, mrays_s=n / total_ms / 1e3, b_ray=bc, roofline_mrays_s=HBM * 1e3 / bc, frac_canonical_hbm=(n / total_ms / 1e3) / (HBM * 1e3 / bc), build_ms=bms, cpu_mrays_s=len(sub) / cpu_s / 1e6, cpu_threads=o.num_threads, cpu_build_ms=obms, intercepted=float(1 - hist[0].item() / n), parity="bit-identical (1M-ray subsample)" if np.array_equal(subc, refc) else "MISMATCH")
    del s, o

if "c4" in want:
    v, t = syn.plot_mesh(4, 25, 1_000_000, 40.0); s, bms = build(v, t); st = s.stats()
    rays = gen_grid(s, syn.sun_direction(60, 30), 4000, 4000); n = rays.shape[0]
    out = [torch.empty(n, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, dtype=torch.uint32, device="cuda"), torch.empty(n, 2, device="cuda"), torch.empty(n, 3, device="cuda")]
    ms = gpu_time(lambda: _lib.check(L.qsmrt_cast_rays_2d(s._h, P(rays), 4000, 4000, *[P(x) for x in out], None)))
    o, obms = oracle_scene(v, t); sub = rays[::64].cpu().numpy(); t0 = time.perf_counter(); ref = o.cast_rays(sub, 1); cpu_s = time.perf_counter() - t0; nn, nt = o.last_counters
    b = 24 + 32 + 32 * nn / len(sub) + 48 * nt / len(sub)
    same = bool(np.array_equal(out[2][::64].cpu().numpy(), ref["primitive_ids"]) and np.array_equal(out[0][::64].cpu().numpy(), ref["t_hit"]))
    report(config="C4 plot (50M triangles) build + cast_rays 16M rays", triangles=int(t.shape[0]), rays=n, mrays_s=n / ms / 1e3, b_ray=b, roofline_mrays_s=HBM * 1e3 / b, frac_canonical_hbm=(n / ms / 1e3) / (HBM * 1e3 / b), build_ms=bms, sort_ms=st["sort_ms"], build_gbs=460.0 * t.shape[0] / bms / 1e6, bvh_height=st["bvh_height"], quantised_nodes=st["quantised_nodes"], cpu_mrays_s=len(sub) / cpu_s / 1e6, cpu_threads=o.num_threads, cpu_build_ms=obms, parity="bit-identical (250k-ray subsample)" if same else "MISMATCH")
