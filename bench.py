#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: Mrays/s of cast_rays on the 2M-triangle
canopy mesh (config C2: 16M parallel sun rays per solar angle, 64-angle
hemisphere sweep) at 1/2/4/8 B200, plus the LBVH build time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one solar angle: cast_rays over 16 777 216 rays (+ the per-triangle
exposure accumulation the sweep keeps on the device).  N > 1 (torchrun, one
rank per GPU): the mesh is broadcast once over NCCL, every rank builds the
same LBVH and takes its own angles of the sweep (weak scaling, no data-path
collective); one all-reduce of the per-triangle exposure closes the timed
region.  Prints ONE JSON line (rank 0).

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


def angles_for(rank, world, nsteps):
    """Solar angles of one rank.  The 8 x 8 sweep is sharded by AZIMUTH: step s = 8 j + e casts elevation e at azimuth
    index (rank * (8 / world) + j + e) mod 8.  The ranks of a step are 360 / world degrees apart -- for 2 and 4 ranks
    the same view of the (x/y symmetric) canopy and its ray grid, i.e. equal work per step -- and every rank walks
    through all azimuths as the elevation changes, so axis-aligned and diagonal views (a few per cent apart in cost)
    are mixed evenly in any run of steps.  N in {1, 2, 4, 8} ranks partition the 64 angles exactly."""
    from pyqsm_b200 import synthetic as syn
    sweep = syn.hemisphere_sweep()                      # index = elevation * 8 + azimuth
    stride = max(1, 8 // world)
    out = []
    for s in range(nsteps):
        e, j = s % 8, s // 8
        out.append(sweep[e * 8 + (rank * stride + j + e) % 8])
    return out


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
    angs = angles_for(0, 1, args.warmup + args.steps)
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
    sample = f"every {stride}th ray of each angle ({n} rays/step), canonical LBVH, OpenMP"
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

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
    angs = angles_for(rank, world, nsteps)
    P = lambda x: C.c_void_p(x.data_ptr())
    F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    # inputs resident in HBM before the timed region: one ray buffer per step
    nbuf = min(nsteps, 16)
    rays = [torch.empty(n, 6, dtype=torch.float32, device=dev) for _ in range(nbuf)]
    for s in range(nbuf):
        g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(*angs[s]), GRID, GRID)
        _lib.check(L.qsmrt_gen_parallel_rays(P(rays[s]), GRID, GRID, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), stream))
    t_hit = torch.empty(n, dtype=torch.float32, device=dev)
    gid = torch.empty(n, dtype=torch.uint32, device=dev)
    pid = torch.empty(n, dtype=torch.uint32, device=dev)
    uv = torch.empty(n, 2, dtype=torch.float32, device=dev)
    nrm = torch.empty(n, 3, dtype=torch.float32, device=dev)
    exposure = torch.zeros(ntri, dtype=torch.int32, device=dev)

    def step(s):
        r = rays[s % nbuf]
        _lib.check(L.qsmrt_cast_rays_2d(scene._h, P(r), GRID, GRID, P(t_hit), P(gid), P(pid), P(uv), P(nrm), stream))
        _lib.check(L.qsmrt_accumulate_hits(scene._h, P(gid), P(pid), n, P(exposure), stream))

    for s in range(args.warmup):
        step(s)
    if world > 1:
        # warm the collective the timed region ends with: NCCL sets up an all-reduce's channels on its first call
        dist.all_reduce(torch.zeros_like(exposure))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    torch.cuda.synchronize()
    ev[0].record()
    for s in range(args.steps):
        ev[1 + 2 * s].record()
        _lib.check(L.qsmrt_cast_rays_2d(scene._h, P(rays[(args.warmup + s) % nbuf]), GRID, GRID, P(t_hit), P(gid), P(pid), P(uv), P(nrm), stream))
        ev[2 + 2 * s].record()
        _lib.check(L.qsmrt_accumulate_hits(scene._h, P(gid), P(pid), n, P(exposure), stream))
    if world > 1:
        dist.all_reduce(exposure)                       # the sweep's only exchange: per-triangle exposure
    ev[-1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[1 + 2 * s].elapsed_time(ev[2 + 2 * s]) for s in range(args.steps)]
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6
    hit_fraction = float(torch.isfinite(t_hit).float().mean().item())

    # ---- e2e: the public API with HOST buffers (pinned), copies inside the timed region
    e2e_steps = max(2, min(args.steps, 5))
    host_scene = RaycastingScene(device=dev)            # CPU results, like Open3D
    host_scene.add_triangles(v, t.view(torch.uint32))
    host_scene.commit()
    host_rays = [rays[s % nbuf].cpu().pin_memory() for s in range(2)]
    # warm-up: staging buffers, and TWO result sets so torch's pinned-host pool holds both
    # generations (the previous result is still referenced while the next call allocates)
    warm = [host_scene.cast_rays(host_rays[0]), host_scene.cast_rays(host_rays[1])]
    ans = host_scene.cast_rays(host_rays[0])
    del warm
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        ans = host_scene.cast_rays(host_rays[s % 2])
        _ = float(ans["t_hit"][0])                      # results are host tensors already
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tm = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    e2e_val = world * n * e2e_steps / float(tm.item()) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_cast_rays), BASELINE.md section 4
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    timed_angles = angs[args.warmup:]
    b_ray, nn, nt = b_ray_for(timed_angles)
    k_ms = float(np.mean(kern_ms))
    achieved = b_ray * n / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "k_trace5<0> (cast_rays, persistent traversal)", "kernel_ms": k_ms, "bytes_per_ray": b_ray,
                "n_node": nn, "n_tri": nt, "peak_source": peak_src,
                "roofline_mrays_s": peak * 1e9 / b_ray / 1e6, "kernel_mrays_s": n / (k_ms * 1e-3) / 1e6}
    prof = os.path.join(ROOT, "profiles", "r01_cast_rays_summary.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline["note"] = ("frac > 1 is expected here: the denominator charges every node/triangle fetch of the canonical "
                        "traversal to HBM, but ncu shows DRAM traffic ~= rays + results only (traffic field) -- the BVH is "
                        "served from L1/L2 -- and the kernel is ALU-pipe / issue bound (sm__inst_executed_pipe_alu ~67-71%, "
                        "issue ~69-71%, L1 data pipe ~61% of peak; profiles/r01_cast_rays_summary.json)")

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
               "sample": f"every 2nd ray of one angle ({sample.shape[0]} rays), oracle canonical LBVH, OpenMP",
               "build_ms": cpu_build_ms}
        # parity spot check on the way (checker only): the e2e answer vs the oracle on that sample
        ans = host_scene.cast_rays(host_rays[0])
        same = bool(np.array_equal(ans["primitive_ids"].numpy()[::2], ref["primitive_ids"]) and
                    np.array_equal(ans["t_hit"].numpy()[::2], ref["t_hit"]))
        cpu["parity_on_sample"] = "bit-identical" if same else "MISMATCH"

    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "triangles": ntri, "rays_per_step": n, "rays_per_step_per_gpu": n,
                   "parallelism": f"replicated scene, angles sharded over {world} GPU(s)",
                   "l2": "no explicit flush: each step reads 384 MB of rays and writes 512 MB of results (> 126 MB L2), "
                         "a different ray buffer every step; the BVH (scene) stays warm across the sweep by design",
                   "hit_fraction_last_step": hit_fraction},
        "build_ms": float(min(builds)), "build_ms_all": [float(b) for b in builds], "sort_ms": float(st["sort_ms"]),
        "bvh": {"nodes": int(st["num_bvh_nodes"]), "leaves": int(st["num_bvh_leaves"]), "bytes": int(st["bvh_bytes"])},
        "mesh_broadcast_ms": bcast_ms,
        "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 32,
                "steps": e2e_steps, "api": "RaycastingScene.cast_rays(pinned host rays) -> host tensors"},
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
