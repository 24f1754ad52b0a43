"""Renders DESIGN.md section 6 from the committed measurement files in profiles/ (r02_bench_n*.json, r02_bench_ref.json,
r02_configs.json, r02_c4_n*.json), so the tables cannot drift from the JSON the driver-style runs produced.
    python tools/make_measured_md.py > /tmp/measured.md"""
import json, os
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path): return None
    lines = [l for l in open(path).read().splitlines() if l.strip().startswith("{")]
    return [json.loads(l) for l in lines]
b = {n: (load(f"r02_bench_n{n}.json") or [None])[-1] for n in (1, 2, 4, 8)}
ref = (load("r02_bench_ref.json") or [None])[-1]
d = b[1]
out = []
if d:
    r = d["roofline"]
    out += ["`python bench.py --steps 20 --warmup 5` (N = 1; `profiles/r02_bench_n1.json`), SM clock %.0f MHz (NVML, %d samples inside the timed" % (d["clocks"]["sm_mhz"], d["clocks"]["samples"]),
            "region), throttle reasons: %s:" % (d["clocks"]["reasons"] or "none"), "",
            "| Quantity | Value |", "|---|---|",
            "| `value` — cast_rays, C2 (2M tris, 16M rays/step, all five outputs), rays + results resident in HBM, incl. exposure accumulation | **%.2f Grays/s** (%.3f ms/step; round 1: 4.99) |" % (d["value"] / 1e3, d["ms_per_step"]),
            "| dominant kernel `k_trace5<0,false,true>` alone (CUDA events) | %.3f ms → %.0f Mrays/s mean over the timed angles |" % (r["kernel_ms"], r["kernel_mrays_s"]),
            "| `roofline` — bound: %s | frac **%.2f** = %.0f of %.0f %s useful lane work; raw issue fraction %.2f at %.1f of 32 lanes, %.0f warp-instructions per ray |" % (r["bound"], r["frac"], r["achieved"], r["peak"], r["unit"], r.get("issue_frac", float("nan")), r.get("lanes_per_inst", float("nan")), r.get("warp_inst_per_ray", float("nan"))),
            "| … compulsory DRAM view | %.0f GB/s = %.3f of the measured %.0f GB/s (`traffic` %.2f GB per launch from ncu) |" % (r["hbm"]["achieved"], r["hbm"]["frac"], r["hbm"]["peak"], (r["traffic"] or 0) / 1e9),
            "| … L2 → SM view | %.0f GB/s of fetches (%.0f B/ray from the kernel's own counters) = %.2f of the %.0f GB/s an in-tree 256-bit read sweep reaches |" % (r["l2"]["achieved"], r["l2"]["bytes_per_ray"], r["l2"]["frac"], r["l2"]["peak"]) if "l2" in r else "",
            "| … `frac_canonical_hbm` (SURVEY §8d bytes, kept for continuity) | %.2f |" % r["frac_canonical_hbm"],
            "| `e2e` — `cast_rays(pinned host rays, outputs=\"all\")` → five host tensors, copies in the timed region | %.0f Mrays/s = %.0f %% of this box's measured duplex PCIe ceiling (%.0f / %.0f GB/s H2D / D2H alone, %.0f GB/s both ways at once) |" % (d["e2e"]["value"], 100 * d["e2e"]["frac_of_link_ceiling"], d["host_link_gbs"]["h2d"][0], d["host_link_gbs"]["d2h"][0], d["host_link_gbs"]["duplex"][0]),
            "| `e2e_ref_pattern` — `cast_rays(pinned host rays)`: `t_hit` + `primitive_ids` eager (8 B/ray back), the rest lazy | **%.0f Mrays/s** = %.0f %% of the ceiling for those bytes (round 1, 32 B/ray always: 1395) |" % (d["e2e_ref_pattern"]["value"], 100 * d["e2e_ref_pattern"]["frac_of_link_ceiling"]),
            "| `fused_sun_sweep` — `environment.sun_exposure`, 64 angles × 16M rays in one launch, per-triangle + per-vertex counts, wall clock | **%.0f Mrays/s** (%.3f s; round 1, 64 launches: 4426) |" % (d["fused_sun_sweep"]["mrays_s"], d["fused_sun_sweep"]["seconds"]),
            "| `fused_sky` — C5: 1M leaf vertices × 1000 hemisphere directions, wall clock | **%.0f Mrays/s** (%.3f s; round 1: 2315) |" % (d["fused_sky"]["mrays_s"], d["fused_sky"]["seconds"]),
            "| LBVH build, 2M triangles (best of 4 commits) | **%.3f ms** (sort %.3f ms); B_tri = 460 B → %.2f TB/s = %.0f %% of the HBM peak (round 1: 0.510 ms) |" % (d["build_ms"], d["sort_ms"], 460 * 2e6 / d["build_ms"] / 1e9, 100 * 460 * 2e6 / d["build_ms"] / 1e9 / 6.552),
            "| CPU baseline — **the repo's scalar LBVH port of Open3D's semantics, not Embree** (%d host threads, 8M-ray sample) | %.1f Mrays/s, CPU build %.0f ms; `--impl reference`: %s Mrays/s; parity on the sample: %s |" % (d["cpu_baseline"]["cores"], d["cpu_baseline"]["value"], d["cpu_baseline"]["build_ms"], ("%.1f" % ref["value"]) if ref else "n/a", d["cpu_baseline"]["parity_on_sample"]),
            ""]
rows = [("N", "value Mrays/s", "ms/step", "% of N × one GPU", "per-rank compute ms (min–max)", "all-reduce ms (max)", "e2e / e2e_ref Mrays/s", "host duplex GB/s (sum)", "fused sun / sky Mrays/s", "parity")]
for n in (1, 2, 4, 8):
    x = b[n]
    if not x: continue
    pr = x["per_rank"]
    rows.append((str(n), "%.0f" % x["value"], "%.3f" % x["ms_per_step"], "%.1f" % (100 * x["value"] / (n * b[1]["value"])) if b[1] else "", "%.2f–%.2f" % (min(pr["compute_ms"]), max(pr["compute_ms"])),
                 "%.2f" % max(pr["allreduce_ms"]), "%.0f / %.0f" % (x["e2e"]["value"], x["e2e_ref_pattern"]["value"]), "%.0f" % x["host_link_gbs"]["duplex_sum"],
                 "%.0f / %.0f" % (x["fused_sun_sweep"]["mrays_s"], x["fused_sky"]["mrays_s"]), x["multi_gpu_parity"]))
out += ["Scaling (`gpurun --gpus N`, torchrun, weak: one 16M-ray angle per step and rank; `profiles/r02_bench_n*.json`):", ""]
out += ["| " + " | ".join(r) + " |" for r in rows[:1]] + ["|" + "---|" * len(rows[0])] + ["| " + " | ".join(r) + " |" for r in rows[1:]] + [""]
c4 = [(n, (load(f"r02_c4_n{n}.json") or [None])[-1]) for n in (2, 4, 8)]
c4 = [(n, x) for n, x in c4 if x]
if c4:
    out += ["Config 4 on N GPUs (`tests/measure/run_c4_multi.py`: 50M triangles, NCCL mesh broadcast of 1.8 GB, identical build on every rank, 16M rays dealt in row blocks, gather on rank 0):", "",
            "| N | broadcast ms (GB/s) | build ms per rank | cast ms per rank (max) | Mrays/s | gather ms | parity |", "|---|---|---|---|---|---|---|"]
    for n, x in c4:
        out.append("| %d | %.1f (%.0f) | %.2f–%.2f | %.2f | %.0f | %.1f | %s |" % (n, x["mesh_broadcast_ms"], x["broadcast_gbs"], min(x["build_ms_per_rank"]), max(x["build_ms_per_rank"]), max(x["cast_ms_per_rank"]), x["mrays_s"], x["gather_ms"], x["multi_gpu_parity"]))
    out.append("")
cfg = load("r02_configs.json")
if cfg:
    out += ["All five BASELINE configurations at full size on one B200 (`tests/measure/run_configs.py` → `profiles/r02_configs.json`; the same launches are",
            "checked by `pytest -m gpu`: `test_c1_*`, `test_c2_full_size_properties`, `test_c3_rain_count_full`, `test_c4_plot_build_and_cast`, `test_c5_sky_full`).",
            "CPU figures: the scalar oracle port on the box's host threads, on the stated subsample.", "",
            "| Config | Triangles | Rays | Mrays/s | build ms | CPU port Mrays/s (threads) | parity |", "|---|---|---|---|---|---|---|"]
    for x in cfg:
        out.append("| %s | %s | %s | **%.0f** | %s | %s | %s |" % (x["config"], "{:,}".format(x["triangles"]) if "triangles" in x else "—", "{:,}".format(x["rays"]), x["mrays_s"],
                   ("%.2f" % x["build_ms"]) if "build_ms" in x else "—", ("%.1f (%d)" % (x["cpu_mrays_s"], x.get("cpu_threads", 16))) if "cpu_mrays_s" in x else "—", x.get("parity", "—")))
    out.append("")
print("\n".join(out))
