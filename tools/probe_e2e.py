"""End-to-end cast_rays (pinned host rays in, host results out) on C2 for several pipeline chunk sizes
(scene options host_chunk = rays per stage of the three-stream host pipe, host_ramp = short stages at both ends).  One line per setting; not product code."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn
v, t = syn.canopy_mesh(2, 1_000_000)
s = RaycastingScene(); s.add_triangles(v, t); s.commit()
st = s.stats(); G = 4000; n = G * G
rays = []
for el, az in ((50, 0), (30, 135)):
    g = syn.parallel_ray_grid(np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64), syn.sun_direction(el, az), G, G)
    rays.append(torch.from_numpy(syn.materialize_grid(*g, G, G).reshape(-1, 6)).pin_memory())
for chunk, ramp in ((1 << 20, 0), (1 << 20, 1), (1 << 21, 0), (1 << 21, 1), (1 << 22, 1), (3 << 19, 1), (1 << 20, 0), (1 << 21, 1)):
    s.set_option("host_chunk", chunk); s.set_option("host_ramp", ramp)
    for outputs, name in (("all", "all five"), (None, "t_hit+prim")):
        warm = [s.cast_rays(rays[0], outputs=outputs), s.cast_rays(rays[1], outputs=outputs)]
        a = s.cast_rays(rays[0], outputs=outputs); del warm
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for k in range(5):
            a = s.cast_rays(rays[k % 2], outputs=outputs); _ = float(a["t_hit"].reshape(-1)[0])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"chunk {chunk:8d} rays ramp {ramp}, {name:10s}: {5 * n / dt / 1e6:7.1f} Mrays/s", flush=True)
