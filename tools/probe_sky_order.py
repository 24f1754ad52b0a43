"""Does the ORDER of the query points matter for the sky Monte-Carlo?  The same 250k leaf vertices in mesh order
(random in space) and sorted along a Morton curve.  Not product code."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, environment as env
v, t = syn.canopy_mesh(2, 1_000_000)
s = RaycastingScene(output_device="cuda"); s.add_triangles(v, t); s.commit()
tri = t.reshape(-1, 2, 3)[:, 0]; p0 = v[tri[:, 0]]; nrm = np.cross(v[tri[:, 1]] - p0, v[tri[:, 2]] - p0); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
step = int(sys.argv[1]) if len(sys.argv) > 1 else 4
p0, nrm = p0[::step].copy(), nrm[::step].astype(np.float32)
def morton(p):
    q = ((p - p.min(0)) / (p.max(0) - p.min(0)) * 1023).astype(np.uint64)
    def spread(x):
        x = (x | (x << 16)) & 0x030000FF; x = (x | (x << 8)) & 0x0300F00F; x = (x | (x << 4)) & 0x030C30C3; x = (x | (x << 2)) & 0x09249249; return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
order = np.argsort(morton(p0), kind="stable")
ident = np.arange(len(p0))
for name, idx, lib_sort in (("mesh order, library sort off", ident, 0), ("mesh order, library sort on", ident, 1), ("pre-sorted, library sort off", order, 0),
                            ("mesh order, library sort off", ident, 0), ("mesh order, library sort on", ident, 1), ("pre-sorted, library sort off", order, 0)):
    s.set_option("point_order", lib_sort)
    pts, nd = torch.from_numpy(p0[idx].copy()).cuda(), torch.from_numpy(nrm[idx].copy()).cuda()
    env.sky_gap_fraction(s, pts, nd, n_dirs=8); torch.cuda.synchronize()
    env.sky_gap_fraction(s, pts[:10000], nd[:10000], n_dirs=100); torch.cuda.synchronize()
    t0 = time.perf_counter(); g = env.sky_gap_fraction(s, pts, nd, n_dirs=1000); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"sky {len(p0)} pts x 1000 dirs, {name}: {pts.shape[0]*1e3/dt/1e6:.0f} Mr/s  mean gap {float(g.mean()):.4f}", flush=True)
