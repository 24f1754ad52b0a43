"""Ad-hoc: where does the time of `rcs(); add_triangles; commit; del` go (50k-triangle tree)?"""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn
v, t = syn.qsm_tree_mesh(seed=1)
vd, td = torch.from_numpy(v).cuda(), torch.from_numpy(t.view(np.int32)).cuda().view(torch.uint32)
T = {k: [] for k in ("create", "add_host", "commit", "del", "add_dev", "commit_dev")}
for rep in range(25):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); s = RaycastingScene(); t1 = time.perf_counter()
    s.add_triangles(v, t); t2 = time.perf_counter()
    s.commit(); t3 = time.perf_counter()
    del s; t4 = time.perf_counter()
    s = RaycastingScene()
    t5 = time.perf_counter(); s.add_triangles(vd, td); t6 = time.perf_counter()
    s.commit(); t7 = time.perf_counter()
    del s
    for k, x in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t6 - t5, t7 - t6)): T[k].append(x * 1e3)
for k, x in T.items(): print(f"{k:12s} median {np.median(x[5:]):8.3f} ms  min {min(x):8.3f}")
