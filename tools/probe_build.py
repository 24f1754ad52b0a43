"""Ad-hoc: LBVH build time, classic vs onesweep sort (not product code)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
L = _lib.load()
for ncan in ([int(a) for a in sys.argv[1:]] or [2, 10, 50]):
    if ncan == 2: v, t = syn.canopy_mesh(2, 1_000_000)
    else: v, t = syn.plot_mesh(3, ncan // 2, 1_000_000, 14.0 if ncan == 10 else 40.0)
    vd, td = torch.from_numpy(v).cuda(), torch.from_numpy(t.view(np.int32)).cuda().view(torch.uint32)
    ref = None
    for variant in (0, 1):
        bs = []
        for i in range(4):
            s = RaycastingScene(output_device="cuda"); s.set_option("sort_variant", variant); s.add_triangles(vd, td); bs.append(s.commit()); st = s.stats()
            if i == 3:
                n = t.shape[0]; keys = np.empty(n, np.uint64); order = np.empty(n, np.uint32)
                _lib.check(L.qsmrt_debug_get_build(s._h, keys.ctypes.data_as(C.c_void_p), order.ctypes.data_as(C.c_void_p), None))
                if ref is None: ref = (keys, order)
                same = np.array_equal(ref[0], keys) and np.array_equal(ref[1], order) and bool(np.all(keys[1:] >= keys[:-1]))
            del s
        print(f"{t.shape[0]/1e6:.0f}M tris sort variant {variant}: build {min(bs):.3f} ms sort {st['sort_ms']:.3f} ms  sorted+identical {same}", flush=True)
    del vd, td
