"""Writes baseline/canonical_counters.json: mean node fetches / triangle tests per
ray of the CANONICAL traversal (the oracle's 63-bit-Morton Karras LBVH, one
triangle per leaf, 32-byte nodes, near-child-first, closest-hit shrinking) on a
fixed subsample (every 16th ray) of each benchmark ray set.  These define the
algorithmic bytes per ray of the roofline:
    B_ray = 24 + OUT + 32 * N_node + 48 * N_tri
They depend only on (mesh, rays), never on the GPU kernel.

    python tests/measure/make_canonical_counters.py            # ~2 min on 8 threads
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from pyqsm_b200 import synthetic as syn

out = {"definition": "mean per-ray fetches of the oracle's canonical LBVH traversal on every 16th ray; "
                     "B_ray = 24 + OUT + 32*N_node + 48*N_tri", "configs": {}}

def run(name, v, t, nu, nv, angles, stride, op="cast"):
    s = oracle.OracleScene(); s.add_triangles(v, t); s.commit()
    lo, hi = v.min(0), v.max(0)
    rows = []
    for (el, az) in angles:
        rays = syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(el, az), nu, nv), nu, nv)[::stride]
        t0 = time.time()
        if op == "cast":
            a = s.cast_rays(rays, 1); hit = float(np.isfinite(a["t_hit"]).mean())
        else:
            c = s.count_intersections(rays, 1); hit = float((c > 0).mean())
        dt = time.time() - t0
        nn, nt = s.last_counters
        rows.append(dict(elevation=el, azimuth=az, n_node=nn / len(rays), n_tri=nt / len(rays), hit_fraction=hit,
                         sample_rays=len(rays), oracle_mrays_s=len(rays) / dt / 1e6))
        print(name, rows[-1], flush=True)
    return rows

v, t = syn.qsm_tree_mesh(1)
out["configs"]["c1_qsm_tree_50k_cast"] = run("c1", v, t, 1000, 1000, [(45.0, 135.0)], 1)
out["configs"]["c1_qsm_tree_50k_count"] = run("c1", v, t, 1000, 1000, [(45.0, 135.0)], 1, "count")
v, t = syn.canopy_mesh(2)
out["configs"]["c2_canopy_2m_cast"] = run("c2", v, t, 4000, 4000, syn.hemisphere_sweep(), 16)
for k, rows in out["configs"].items():
    nn = float(np.mean([r["n_node"] for r in rows])); nt = float(np.mean([r["n_tri"] for r in rows]))
    OUT = 4 if k.endswith("count") else 32
    out.setdefault("summary", {})[k] = dict(n_node=nn, n_tri=nt, out_bytes=OUT, b_ray=24 + OUT + 32 * nn + 48 * nt)
os.makedirs(os.path.join(os.path.dirname(__file__), "..", "..", "baseline"), exist_ok=True)
with open(os.path.join(os.path.dirname(__file__), "..", "..", "baseline", "canonical_counters.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out["summary"], indent=1))
