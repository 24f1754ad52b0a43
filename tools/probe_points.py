"""Ad-hoc: latency split of the point queries at the reference's sizes (256 points, 64^3 grid) on the C1 tree."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn
v, t = syn.qsm_tree_mesh(seed=1)
s = RaycastingScene(); s.add_triangles(v, t); s.commit()
sd = RaycastingScene(output_device="cuda"); sd.add_triangles(v, t); sd.commit()
lo, hi = v.min(0), v.max(0)
def bench(name, f, reps=30):
    for _ in range(3): f()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:52s} median {np.median(ts):8.3f} ms  min {min(ts):8.3f} ms", flush=True)
for n in (1, 256, 4096, 65536):
    q = np.random.default_rng(0).uniform(lo, hi, size=(n, 3)).astype(np.float32)
    qd = torch.from_numpy(q).cuda()
    bench(f"n={n}: compute_distance (host in/out)", lambda: s.compute_distance(q))
    bench(f"n={n}: compute_distance (device in/out)", lambda: sd.compute_distance(qd))
    bench(f"n={n}: compute_occupancy (device)", lambda: sd.compute_occupancy(qd))
    bench(f"n={n}: compute_signed_distance (device)", lambda: sd.compute_signed_distance(qd))
    bench(f"n={n}: compute_closest_points (device)", lambda: sd.compute_closest_points(qd))
