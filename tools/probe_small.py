"""Ad-hoc: latency of the reference's own call shapes (ray_casting.py: 640x480 and 1280x950 pinhole images, 100-ray
list_intersections, 1M-ray C1 batch) through the Python API with host tensors in and out."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn
v, t = syn.qsm_tree_mesh(seed=1)
s = RaycastingScene(); s.add_triangles(v, t); s.commit()
c = (v.min(0) + v.max(0)) / 2
def bench(name, f, reps=30):
    for _ in range(3): f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:46s} median {np.median(ts):8.3f} ms  min {min(ts):8.3f} ms", flush=True)
for (w, h) in ((640, 480), (1280, 950)):
    bench(f"create_rays_pinhole {w}x{h}", lambda: RaycastingScene.create_rays_pinhole(fov_deg=60, center=list(c), eye=[c[0], c[1] - 15, c[2]], up=[0, 0, 1], width_px=w, height_px=h))
    rays = RaycastingScene.create_rays_pinhole(fov_deg=60, center=list(c), eye=[c[0], c[1] - 15, c[2]], up=[0, 0, 1], width_px=w, height_px=h)
    print("  rays", type(rays), rays.device, tuple(rays.shape))
    bench(f"cast_rays {w}x{h} host->host ({w*h} rays)", lambda: s.cast_rays(rays))
    rn = rays.numpy()
    bench(f"cast_rays {w}x{h} numpy in", lambda: s.cast_rays(rn))
    bench(f"count_intersections {w}x{h}", lambda: s.count_intersections(rays))
grid = syn.parallel_ray_grid(v.min(0), v.max(0), syn.sun_direction(45, 135), 1000, 1000)
r1 = torch.from_numpy(syn.materialize_grid(*grid, 1000, 1000))
bench("cast_rays C1 1M rays host->host", lambda: s.cast_rays(r1), 10)
bench("cast_rays C1 1M rays pinned host", (lambda rp: (lambda: s.cast_rays(rp)))(r1.pin_memory()), 10)
r100 = r1[::10000].contiguous()
bench("list_intersections 100 rays", lambda: s.list_intersections(r100))
bench("cast_rays 100 rays", lambda: s.cast_rays(r100))
sd = RaycastingScene(output_device="cuda"); sd.add_triangles(v, t); sd.commit()
rd = r1.cuda()
def dev(): sd.cast_rays(rd); torch.cuda.synchronize()
bench("cast_rays C1 1M rays device->device (+sync)", dev, 10)
def build():
    x = RaycastingScene(); x.add_triangles(v, t); x.commit()
bench("scene create + add_triangles(host) + commit 50k tris", build, 10)
# ray_casting.py:241-255: signed distance on 256 random points and on a 64^3 grid
lo, hi = v.min(0), v.max(0)
q256 = np.random.default_rng(0).uniform(lo, hi, size=(256, 3)).astype(np.float32)
xs = [np.linspace(lo[a], hi[a], 64, dtype=np.float32) for a in range(3)]
g64 = np.stack(np.meshgrid(*xs, indexing="ij"), -1)
bench("compute_signed_distance 256 points", lambda: s.compute_signed_distance(q256))
bench("compute_signed_distance 64^3 grid", lambda: s.compute_signed_distance(g64), 10)
bench("compute_distance 64^3 grid", lambda: s.compute_distance(g64), 10)
bench("compute_occupancy 64^3 grid", lambda: s.compute_occupancy(g64), 10)
bench("compute_closest_points 64^3 grid", lambda: s.compute_closest_points(g64), 10)
vc, tc = syn.canopy_mesh(2, 1_000_000)
s2 = RaycastingScene(); s2.add_triangles(vc, tc); s2.commit()
lo, hi = vc.min(0), vc.max(0)
xs = [np.linspace(lo[a], hi[a], 64, dtype=np.float32) for a in range(3)]
g64 = np.stack(np.meshgrid(*xs, indexing="ij"), -1)
bench("canopy 2M tris: compute_signed_distance 64^3 grid", lambda: s2.compute_signed_distance(g64), 5)
r1M = torch.from_numpy(syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(45, 135), 1000, 1000), 1000, 1000))
bench("canopy 2M tris: list_intersections 1M rays", lambda: s2.list_intersections(r1M), 5)
bench("canopy 2M tris: count_intersections 1M rays", lambda: s2.count_intersections(r1M), 5)
