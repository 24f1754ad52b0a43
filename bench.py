#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: Mrays/s of cast_rays on the 2M-triangle
canopy mesh (config C2: 16M parallel sun rays per solar angle, 64-angle
hemisphere sweep) at 1/2/4/8 B200, plus the LBVH build time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one solar angle: cast_rays over 16 000 000 rays (+ the per-triangle
exposure accumulation the sweep keeps on the device).  N > 1 (torchrun, one
rank per GPU): the mesh is broadcast once over NCCL, every rank builds the
same LBVH and casts one whole solar angle (16 000 000 rays) per step; the angle
slots are dealt to the ranks longest-processing-time-first on the committed
per-angle kernel times, so the ranks' totals agree although single angles
differ by +-15 % (weak scaling, no data-path collective); one all-reduce of
the per-triangle exposure closes the timed region, and rank 0 recomputes the
whole timed sweep alone to check the all-reduced result (`multi_gpu_parity`).
Prints ONE JSON line (rank 0).

--impl reference times the CPU path instead: Open3D itself is not installable
in this image (no network), so it is the repo's CPU oracle
(oracle/, canonical LBVH, OpenMP on all host threads) on a bounded sample of
the same rays.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

_JSON_OUT = sys.stdout
GRID = 4000                       # 4000 x 4000 = 16M rays per angle
N_LEAVES = 1_000_000              # 2M triangles
WORKLOAD = "C2: canopy leaf-soup mesh 2M triangles (seed 2), 16M parallel sun rays per angle, 64-angle hemisphere sweep"
METRIC = "Mrays/s cast_rays, 2M-tri canopy mesh at 1/2/4/8 B200; LBVH build ms"


def load_counters():
    with open(os.path.join(ROOT, "baseline", "canonical_counters.json")) as f:
        return json.load(f)


def b_ray_for(angles):
    """Algorithmic bytes per ray averaged over the angles actually timed."""
    rows = load_counters()["configs"]["c2_canopy_2m_cast"]
    key = {(round(r["elevation"], 3), round(r["azimuth"], 3)): r for r in rows}
    sel = [key[(round(e, 3), round(a, 3))] for e, a in angles]
    nn = float(np.mean([r["n_node"] for r in sel]))
    nt = float(np.mean([r["n_tri"] for r in sel]))
    return 24 + 32 + 32 * nn + 48 * nt, nn, nt


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML thread polling every ~2 ms (the region
    is tens of milliseconds, too short for `nvidia-smi -lms`, whose start-up alone takes longer); nvidia-smi one-shot
    as the fallback when NVML cannot be loaded."""
    NAMES = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
             ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.mx, self.h, self.nv = index, [], set(), None, None, None
        self._stop, self._thr = threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, attr in self.NAMES:
                    if r & getattr(nv, attr):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "how": "NVML polled every ~2 ms inside the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": [], "how": "nvidia-smi one-shot after the region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}


def angle_costs():
    """Relative cost of every solar angle of the sweep (kernel ms, index = elevation * 8 + azimuth) from the committed
    per-angle profile of the cast kernel (tools/profile_angles.py); None if the profile is not there."""
    from pyqsm_b200 import synthetic as syn
    path = os.path.join(ROOT, "profiles", "r02_cast_rays_profile.json")
    try:
        rows = json.load(open(path))["angles"]
        key = {(round(r["elevation"], 3), round(r["azimuth"], 3)): float(r["kernel_ms"]) for r in rows}
        return [key[(round(e, 3), round(a, 3))] for e, a in syn.hemisphere_sweep()]
    except Exception:
        return None


def schedule(world, nsteps, warmup=0):
    """Solar angles of every rank: [world][nsteps] (elevation, azimuth).  Whole angles per rank (one 16M-ray grid per
    step and rank: weak scaling, no direction mixing inside a launch, which costs ~7 %: profiles/r02_mix.txt).  The
    world x nsteps angle slots walk the 8 x 8 sweep with stride 27 (coprime with 64: every window of the walk mixes
    elevations and azimuths evenly, and any 64 consecutive slots are the whole sweep).  The first `warmup` steps take
    slots in walk order; the timed slots are dealt longest-processing-time-first on the committed per-angle kernel
    times (angle_costs), every rank getting the same number, so the ranks' totals agree to a fraction of a per cent
    although single angles differ by +-15 %.  Without the profile the deal is round-robin in walk order."""
    from pyqsm_b200 import synthetic as syn
    sweep = syn.hemisphere_sweep()                      # index = elevation * 8 + azimuth
    walk = [(27 * g) % 64 for g in range(world * nsteps)]
    per_rank = [[walk[s * world + r] for s in range(warmup)] for r in range(world)]
    timed = walk[world * warmup:]
    cost = angle_costs()
    k = nsteps - warmup
    if cost is None or world == 1:
        for r in range(world):
            per_rank[r] += timed[r::world]
    else:
        load, mine = [0.0] * world, [[] for _ in range(world)]
        for idx in sorted(timed, key=lambda i: (-cost[i], i)):
            r = min((x for x in range(world) if len(mine[x]) < k), key=lambda x: (load[x], x))
            mine[r].append(idx); load[r] += cost[idx]
        for r in range(world):
            per_rank[r] += sorted(mine[r], key=lambda i: timed.index(i))
    return [[sweep[i] for i in seq] for seq in per_rank]


def angles_for(rank, world, nsteps, warmup=0):
    return schedule(world, nsteps, warmup)[rank]


# --------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from pyqsm_b200 import synthetic as syn
    oracle.build_oracle()
    v, t = syn.canopy_mesh(2, N_LEAVES)
    sc = oracle.OracleScene()
    sc.add_triangles(v, t)
    t0 = time.perf_counter()
    sc.commit()
    build_ms = (time.perf_counter() - t0) * 1e3
    lo, hi = v.min(0), v.max(0)
    stride = 8                                      # 2M of the angle's 16M rays per step
    angs = angles_for(0, 1, args.warmup + args.steps, args.warmup)
    ray_sets = [syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(e, a), GRID, GRID), GRID, GRID)[::stride].copy()
                for e, a in angs[: min(len(angs), 4)]]
    n = ray_sets[0].shape[0]
    for s in range(args.warmup):
        sc.cast_rays(ray_sets[s % len(ray_sets)], 1)
    t0 = time.perf_counter()
    for s in range(args.steps):
        sc.cast_rays(ray_sets[(args.warmup + s) % len(ray_sets)], 1)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt / 1e6
    sample = (f"every {stride}th ray of each angle ({n} rays/step); the repo's scalar LBVH port of Open3D's semantics (oracle/, "
              "OpenMP) -- NOT Embree, whose SIMD BVH is roughly an order of magnitude faster per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "triangles": int(t.shape[0]), "rays_per_step": n,
                   "note": "Open3D/Embree is not installable in this image (no network); the CPU arm is the repo's "
                           "oracle port of its semantics"},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": sc.num_threads, "kind": "port", "sample": sample,
                         "build_ms": build_ms},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _JSON_OUT.write(json.dumps(line) + "\n"); _JSON_OUT.flush()


# --------------------------------------------------------------------- our arm
def fill_rays(L, buf, lo, hi, angle, stream):
    """The 16M parallel rays of one solar angle, generated on the device into `buf`."""
    import ctypes as C
    from pyqsm_b200 import synthetic as syn, _lib
    F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
    g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(*angle), GRID, GRID)
    _lib.check(L.qsmrt_gen_parallel_rays(C.c_void_p(buf.data_ptr()), GRID, GRID, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), stream))


def measure_read_gbs(L, dev, mbytes, reps):
    """GB/s of qsmrt_util_read_sweep over a buffer of `mbytes` MB (256-bit loads, 8 CTAs / SM)."""
    import ctypes as C
    import torch
    buf = torch.zeros(mbytes << 20, dtype=torch.uint8, device=dev)
    sink = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(4):
        L.qsmrt_util_read_sweep(C.c_void_p(buf.data_ptr()), buf.numel(), 1, C.c_void_p(sink.data_ptr()), st)     # warm (L2 fill)
        e0.record()
        L.qsmrt_util_read_sweep(C.c_void_p(buf.data_ptr()), buf.numel(), reps, C.c_void_p(sink.data_ptr()), st)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, buf.numel() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def measure_host_link(dev, h2d_bytes, d2h_bytes, reps=3):
    """Ceiling of the end-to-end path on this box: one cudaMemcpyAsync per direction from / to pinned host memory,
    both directions concurrently on two streams (what the three-stream pipe of cast_rays_host can at best overlap)."""
    import torch
    src_h = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    dst_d = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    src_d = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    dst_h = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {}
    for name, do_in, do_out in (("h2d_gbs", True, False), ("d2h_gbs", False, True), ("both", True, True)):
        best = 1e30
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if do_in:
                with torch.cuda.stream(s1):
                    dst_d.copy_(src_h, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    dst_h.copy_(src_d, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        if name == "both":
            out["duplex_s"] = best
            out["duplex_gbs"] = (h2d_bytes + d2h_bytes) / best / 1e9
        else:
            out[name] = (h2d_bytes if do_in else d2h_bytes) / best / 1e9
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyqsm_b200 import RaycastingScene, synthetic as syn, environment as env, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def gather_f64(vals):
        """[world][len(vals)] float64 on every rank."""
        x = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world == 1:
            return x[None].cpu().numpy()
        out = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(out, x)
        return torch.stack(out).cpu().numpy()

    # ---- scene: rank 0 makes the mesh, NCCL broadcast, every rank builds the same LBVH
    from pyqsm_b200.distributed import broadcast_mesh
    if rank == 0:
        v_np, t_np = syn.canopy_mesh(2, N_LEAVES)
        v = torch.from_numpy(v_np).to(dev)
        t = torch.from_numpy(t_np.view(np.int32)).to(dev)
    else:
        v = t = None
    bcast_ms = 0.0
    if world > 1:
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        v, t = broadcast_mesh(v, t, src=0, device=dev)
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
    L = _lib.load()
    scene = RaycastingScene(device=dev, output_device=dev)
    scene.add_triangles(v, t.view(torch.uint32))
    builds = []
    for _ in range(3):                                  # build time: best of 3 fresh commits
        s2 = RaycastingScene(device=dev, output_device=dev)
        s2.add_triangles(v, t.view(torch.uint32))
        builds.append(s2.commit())
        del s2
    build_ms = scene.commit()
    builds.append(build_ms)
    st = scene.stats()
    lo = np.asarray(st["scene_lo"], np.float64)
    hi = np.asarray(st["scene_hi"], np.float64)
    ntri = int(st["num_triangles"])

    n = GRID * GRID
    nsteps = args.warmup + args.steps
    P = lambda x: C.c_void_p(x.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # inputs resident in HBM before the timed region: one 384 MB ray buffer per step (at most the 64 angles of the sweep)
    plan = schedule(world, nsteps, args.warmup)
    angs = plan[rank]
    nbuf = min(nsteps, 64)
    rays = [torch.empty(n, 6, dtype=torch.float32, device=dev) for _ in range(nbuf)]
    for s in range(max(0, nsteps - nbuf), nsteps):
        fill_rays(L, rays[s % nbuf], lo, hi, angs[s], stream)
    t_hit = torch.empty(n, dtype=torch.float32, device=dev)
    gid = torch.empty(n, dtype=torch.uint32, device=dev)
    pid = torch.empty(n, dtype=torch.uint32, device=dev)
    uv = torch.empty(n, 2, dtype=torch.float32, device=dev)
    nrm = torch.empty(n, 3, dtype=torch.float32, device=dev)
    exposure = torch.zeros(ntri, dtype=torch.int32, device=dev)

    def cast(r):
        _lib.check(L.qsmrt_cast_rays_2d(scene._h, P(r), GRID, GRID, P(t_hit), P(gid), P(pid), P(uv), P(nrm), stream))

    def accumulate(into):
        _lib.check(L.qsmrt_accumulate_hits(scene._h, P(gid), P(pid), n, P(into), stream))

    for s in range(args.warmup):
        cast(rays[s % nbuf]); accumulate(exposure)
    exposure.zero_()
    if world > 1:
        # warm the collective the timed region ends with: NCCL sets up an all-reduce's channels on its first call
        dist.all_reduce(torch.zeros_like(exposure))
    clocks = ClockSampler(local)                        # every rank samples its own GPU (NVML start-up happens here, not
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 3)]     # between the barrier and the start)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                                  # all ranks enter the timed region together: the closing all-reduce
    clocks.start()                                      # would otherwise charge a late starter's delay to everyone else
    ev[0].record()
    for s in range(args.steps):
        ev[1 + 2 * s].record()
        cast(rays[(args.warmup + s) % nbuf])
        ev[2 + 2 * s].record()
        accumulate(exposure)
    ev[-2].record()
    if world > 1:
        dist.all_reduce(exposure)                       # the sweep's only exchange: per-triangle exposure
    ev[-1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop()
    total_ms_mine = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[1 + 2 * s].elapsed_time(ev[2 + 2 * s]) for s in range(args.steps)]
    per_rank = gather_f64([total_ms_mine, ev[0].elapsed_time(ev[-2]), ev[-2].elapsed_time(ev[-1]), float(np.mean(kern_ms)),
                           float(np.max(kern_ms)), float(clk["sm_mhz"] or 0.0), float(len(clk["reasons"]))])
    total_ms = float(per_rank[:, 0].max())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6
    hit_fraction = float(torch.isfinite(t_hit).float().mean().item())

    # ---- N > 1: the all-reduced exposure against the same sweep recomputed by rank 0 alone (full angle grids)
    parity = "n/a (single GPU)"
    if world > 1:
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        if rank == 0:
            ref = torch.zeros_like(exposure)
            full = torch.empty(n, 6, dtype=torch.float32, device=dev)
            for r in range(world):
                for s in range(args.warmup, nsteps):
                    fill_rays(L, full, lo, hi, plan[r][s], stream)
                    cast(full); accumulate(ref)
            ok[0] = 1 if torch.equal(ref, exposure) else 0
            del full
        dist.broadcast(ok, 0)
        parity = "bit-identical" if int(ok.item()) == 1 else "MISMATCH"

    # ---- e2e: the public API with HOST buffers (pinned), copies inside the timed region
    e2e_steps = max(2, min(args.steps, 5))
    host_scene = RaycastingScene(device=dev)            # CPU results, like Open3D
    host_scene.add_triangles(v, t.view(torch.uint32))
    host_scene.commit()
    host_rays = [rays[(nsteps - 1 - s) % nbuf].cpu().pin_memory() for s in range(2)]

    def e2e_run(outputs, touch):
        # warm-up: staging buffers, and TWO result sets so torch's pinned-host pool holds both
        # generations (the previous result is still referenced while the next call allocates)
        warm = [host_scene.cast_rays(host_rays[0], outputs=outputs), host_scene.cast_rays(host_rays[1], outputs=outputs)]
        ans = host_scene.cast_rays(host_rays[0], outputs=outputs)
        del warm
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            ans = host_scene.cast_rays(host_rays[s % 2], outputs=outputs)
            for k in touch:
                _ = float(ans[k].reshape(-1)[0])          # results are host tensors already
        torch.cuda.synchronize()
        return world * n * e2e_steps / float(gather_f64([time.perf_counter() - t0])[:, 0].max()) / 1e6

    e2e_val = e2e_run("all", ("t_hit",))
    e2e_ref = e2e_run(None, ("t_hit", "primitive_ids"))   # the reference's own pattern: t_hit + primitive_ids (8 B/ray back)
    link = measure_host_link(dev, n * 24, n * 32)
    link_ref = measure_host_link(dev, n * 24, n * 8)
    link_all = gather_f64([link["h2d_gbs"], link["d2h_gbs"], link["duplex_gbs"], link["duplex_s"], link_ref["duplex_s"]])

    # ---- the drivers a user of the README's feature calls: the fused sun sweep (64 angles x 16M rays, one launch)
    #      and the sky Monte-Carlo (C5: 1M leaf vertices x 1000 directions); sharded over the ranks when N > 1
    shard = (rank, world) if world > 1 else None
    sweep = syn.hemisphere_sweep()
    env.sun_exposure(scene, sweep[:2 * world], grid=(GRID, GRID), shard=shard)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    fused = env.sun_exposure(scene, sweep, grid=(GRID, GRID), shard=shard, per_vertex=True)
    torch.cuda.synchronize()
    fused_s = float(gather_f64([time.perf_counter() - t0])[:, 0].max())
    tri0 = t.view(-1, 2, 3)[:, 0].long()
    p0, p1, p2 = v[tri0[:, 0]], v[tri0[:, 1]], v[tri0[:, 2]]
    pn = torch.linalg.cross(p1 - p0, p2 - p0)
    pn = pn / pn.norm(dim=1, keepdim=True)
    env.sky_gap_fraction(scene, p0, pn, n_dirs=8 * world, shard=shard)          # warm-up at full size: the point-order scratch is allocated here
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    gap = env.sky_gap_fraction(scene, p0, pn, n_dirs=1000, seed=5, shard=shard)
    torch.cuda.synchronize()
    sky_s = float(gather_f64([time.perf_counter() - t0])[:, 0].max())

    l2_gbs = measure_read_gbs(L, dev, 48, 40) if rank == 0 else 0.0
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_trace5<0>: cast_rays).  ncu shows it is NOT DRAM bound (DRAM ~5 % of
    #      peak: the BVH is served from L1 / L2) but instruction-issue bound, so `frac` is the share of the SM issue
    #      rate doing useful lane work; the DRAM and L2 views and the canonical-bytes figure of SURVEY 8d are beside it
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_hbm, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_hbm, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    timed_angles = angs[args.warmup:]
    b_ray, nn, nt = b_ray_for(timed_angles)
    k_ms = float(np.mean(kern_ms))
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = float(clk["sm_mhz"] or 1965.0)
    prof = {}
    prof_path = os.path.join(ROOT, "profiles", "r02_cast_rays_profile.json")
    if os.path.exists(prof_path):
        try:
            prof = json.load(open(prof_path))
        except Exception:
            prof = {}
    # per-angle table (tools/profile_angles.py): take the angles this run timed (a multi-GPU step mixes its angles evenly)
    key = {(round(x["elevation"], 3), round(x["azimuth"], 3)): x for x in prof.get("angles", [])}
    sel = [key.get((round(e, 3), round(z, 3))) for e, z in timed_angles]
    if sel and all(x is not None for x in sel):
        rpa = float(prof["rays_per_angle"])
        inst_ray = float(np.mean([x["warp_inst"] for x in sel])) / rpa      # ncu smsp__inst_executed.sum / rays
        lanes = float(sum(x["thread_inst"] for x in sel) / sum(x["warp_inst"] for x in sel))
        fetch_b = float(np.mean([32 * x["nodes_per_ray"] + 48 * x["tris_per_ray"] for x in sel]))   # the kernel's own counters
        traffic = float(np.mean([x["dram_bytes"] for x in sel]))
    else:
        inst_ray = lanes = fetch_b = traffic = None
    issue_peak = sms * 4 * sm_mhz * 1e6 / 1e9             # G warp-instructions / s
    roofline = {"bound": "issue", "unit": "Gwarp-inst/s", "peak": issue_peak, "traffic": traffic,
                "kernel": "k_trace5<0> (cast_rays, persistent traversal)", "kernel_ms": k_ms, "kernel_mrays_s": n / (k_ms * 1e-3) / 1e6,
                "peak_source": f"{sms} SMs x 4 schedulers x {sm_mhz:.0f} MHz (NVML, median inside the timed region)"}
    if inst_ray and lanes:
        issued = inst_ray * n / (k_ms * 1e-3) / 1e9
        roofline.update({"achieved": issued * lanes / 32.0, "frac": issued * lanes / 32.0 / issue_peak, "issue_frac": issued / issue_peak,
                         "lanes_per_inst": lanes, "warp_inst_per_ray": inst_ray,
                         "how": "achieved = ncu warp-instructions per ray (profiles/r02_cast_rays_profile.json, same command) x rays "
                                "/ live CUDA-event kernel time x lanes/32: the share of the issue rate doing useful lane work"})
    dram_compulsory = (56.0 * n + float(st["bvh_bytes"])) / (k_ms * 1e-3) / 1e9
    if not (inst_ray and lanes):        # no instruction profile next to this bench.py: fall back to the compulsory-DRAM view
        roofline.update({"bound": "hbm", "unit": "GB/s", "peak": peak_hbm, "achieved": dram_compulsory, "frac": dram_compulsory / peak_hbm,
                         "how": "profiles/r02_cast_rays_profile.json missing: compulsory DRAM bytes / kernel time"})
    roofline["hbm"] = {"achieved": dram_compulsory, "peak": peak_hbm, "unit": "GB/s", "frac": dram_compulsory / peak_hbm, "peak_source": peak_src,
                       "bytes": "24 B ray + 32 B results per ray + the BVH once per launch (compulsory traffic)"}
    if fetch_b:
        l2_ach = fetch_b * n / (k_ms * 1e-3) / 1e9
        roofline["l2"] = {"achieved": l2_ach, "peak": l2_gbs, "unit": "GB/s", "frac": l2_ach / l2_gbs if l2_gbs else None, "bytes_per_ray": fetch_b,
                          "peak_source": "qsmrt_util_read_sweep over a 48 MB buffer (L2 resident), this run"}
    canon = b_ray * n / (k_ms * 1e-3) / 1e9
    roofline["frac_canonical_hbm"] = canon / peak_hbm
    roofline["canonical"] = {"bytes_per_ray": b_ray, "n_node": nn, "n_tri": nt, "achieved_gbs": canon,
                             "note": "SURVEY 8d figure: every node / triangle fetch of the oracle's canonical traversal charged to HBM; kept for "
                                     "continuity, it exceeds 1 because those fetches are served from L1 / L2"}

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample of the same rays
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build_oracle()
        osc = oracle.OracleScene()
        osc.add_triangles(v.cpu().numpy(), t.cpu().numpy().view(np.uint32))
        t0 = time.perf_counter()
        osc.commit()
        cpu_build_ms = (time.perf_counter() - t0) * 1e3
        sample = host_rays[0].numpy()[::2]              # 8M rays of one angle: ~6-10 s on 8 threads
        osc.cast_rays(sample[:200000], 1)
        t0 = time.perf_counter()
        ref = osc.cast_rays(sample, 1)
        dt = time.perf_counter() - t0
        cpu = {"value": sample.shape[0] / dt / 1e6, "unit": "Mrays/s", "cores": osc.num_threads, "kind": "port",
               "sample": f"every 2nd ray of one angle ({sample.shape[0]} rays); the repo's scalar LBVH port of Open3D's semantics "
                         "(oracle/, OpenMP) -- NOT Embree, whose SIMD BVH is roughly an order of magnitude faster per core",
               "build_ms": cpu_build_ms}
        # parity spot check on the way (checker only): the e2e answer vs the oracle on that sample
        ans = host_scene.cast_rays(host_rays[0])
        same = bool(np.array_equal(ans["primitive_ids"].numpy()[::2], ref["primitive_ids"]) and
                    np.array_equal(ans["t_hit"].numpy()[::2], ref["t_hit"]))
        cpu["parity_on_sample"] = "bit-identical" if same else "MISMATCH"

    ceil_all = world * n / float(link_all[:, 3].max()) / 1e6          # rays/s if only the copies of a step had to happen
    ceil_ref = world * n / float(link_all[:, 4].max()) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "triangles": ntri, "rays_per_step": n * world, "rays_per_step_per_gpu": n,
                   "parallelism": f"replicated scene; one whole solar angle (16M rays) per step and rank over {world} rank(s); the timed "
                                  "angle slots (a stride-27 walk of the 8 x 8 sweep) are dealt to the ranks longest-processing-time-"
                                  "first on the committed per-angle kernel times",
                   "l2": "no explicit flush: each step reads 384 MB of rays and writes 512 MB of results (> 126 MB L2), "
                         "a different ray buffer every step; the BVH (scene) stays warm across the sweep by design",
                   "hit_fraction_last_step": hit_fraction},
        "build_ms": float(min(builds)), "build_ms_all": [float(b) for b in builds], "sort_ms": float(st["sort_ms"]),
        "bvh": {"nodes": int(st["num_bvh_nodes"]), "leaves": int(st["num_bvh_leaves"]), "bytes": int(st["bvh_bytes"])},
        "mesh_broadcast_ms": bcast_ms,
        "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 32,
                "steps": e2e_steps, "api": "RaycastingScene.cast_rays(pinned host rays, outputs='all') -> five host tensors",
                "link_ceiling_mrays_s": ceil_all, "frac_of_link_ceiling": e2e_val / ceil_all},
        "e2e_ref_pattern": {"value": e2e_ref, "unit": "Mrays/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 8,
                            "api": "RaycastingScene.cast_rays(pinned host rays) -> t_hit + primitive_ids on the host, the other three keys "
                                   "stay on the GPU until read (what ray_casting.py:279-289,319-322 consumes)",
                            "link_ceiling_mrays_s": ceil_ref, "frac_of_link_ceiling": e2e_ref / ceil_ref},
        "host_link_gbs": {"h2d": [float(x) for x in link_all[:, 0]], "d2h": [float(x) for x in link_all[:, 1]],
                          "duplex": [float(x) for x in link_all[:, 2]], "duplex_sum": float(link_all[:, 2].sum()),
                          "how": "per rank, concurrently on all ranks: one cudaMemcpyAsync of 384 MB in and one of 512 MB out, pinned host memory"},
        "fused_sun_sweep": {"mrays_s": len(sweep) * n / fused_s / 1e6, "rays": len(sweep) * n, "seconds": fused_s, "launches_per_rank": 1,
                            "sunlit_rays": int(fused["counts"].sum().item()), "sunlit_vertices": int((fused["vertex_counts"] > 0).sum().item()),
                            "api": "environment.sun_exposure(64 angles x 16M rays, per_vertex=True): wall clock incl. the all-reduce at N > 1"},
        "fused_sky": {"mrays_s": p0.shape[0] * 1000 / sky_s / 1e6, "rays": int(p0.shape[0]) * 1000, "seconds": sky_s,
                      "mean_gap_fraction": float(gap.mean().item()),
                      "api": "environment.sky_gap_fraction(1M leaf vertices x 1000 directions), directions sharded over the ranks"},
        "multi_gpu_parity": parity,
        "per_rank": {"total_ms": [float(x) for x in per_rank[:, 0]], "compute_ms": [float(x) for x in per_rank[:, 1]],
                     "allreduce_ms": [float(x) for x in per_rank[:, 2]], "kernel_ms_mean": [float(x) for x in per_rank[:, 3]],
                     "kernel_ms_max": [float(x) for x in per_rank[:, 4]], "sm_mhz": [float(x) for x in per_rank[:, 5]],
                     "throttle_reasons": [int(x) for x in per_rank[:, 6]]},
        "gpu_launches": 2 * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
    }
    _JSON_OUT.write(json.dumps(line) + "\n"); _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries the ONE JSON line and nothing else: libraries that print there (NCCL's version banner does) are
    # sent to stderr, the line itself is written to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "ours":
        # launched bare: re-exec under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd, stdout=_JSON_OUT))
    if args.impl == "reference":
        # the CPU arm uses every host thread (torchrun exports OMP_NUM_THREADS=1 to its workers); set before libgomp loads
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
