"""Ad-hoc: host-buffer (e2e) path vs pipeline chunk size (QSMRT_HOST_CHUNK)."""
import os, sys, time, subprocess
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from pyqsm_b200 import RaycastingScene, synthetic as syn
    v, t = syn.canopy_mesh(2, 1_000_000)
    s = RaycastingScene(); s.add_triangles(v, t); s.commit()
    rays = torch.from_numpy(syn.materialize_grid(*syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(40, 30), 4000, 4000), 4000, 4000)).pin_memory()
    keep = [s.cast_rays(rays), s.cast_rays(rays)]; del keep
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); a = s.cast_rays(rays); best = min(best, time.perf_counter() - t0)
    print(f"chunk {os.environ.get('QSMRT_HOST_CHUNK','default')}: {best*1e3:.2f} ms {16e6/best/1e6:.0f} Mrays/s", flush=True)
else:
    for c in ("262144", "524288", "1048576", "2097152", "4194304"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, QSMRT_HOST_CHUNK=c))
