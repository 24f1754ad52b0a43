"""Ad-hoc probe of the host (e2e) path: PCIe bandwidth and chunked pipeline timing."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
n = 16_000_000
h = torch.empty(n, 6, pin_memory=True); d = torch.empty(n, 6, device="cuda")
o = torch.empty(n, 8, pin_memory=True); do = torch.empty(n, 8, device="cuda")
for name, f in (("h2d 384MB", lambda: d.copy_(h, non_blocking=True)), ("d2h 512MB", lambda: o.copy_(do, non_blocking=True))):
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, f"{dt*1e3:.2f} ms", f"{(384e6 if 'h2d' in name else 512e6)/dt/1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): o.copy_(do, non_blocking=True)
torch.cuda.synchronize(); print("duplex both", f"{(time.perf_counter()-t0)*1e3:.2f} ms")
t0 = time.perf_counter(); x = torch.empty(n, 8, pin_memory=True); print("pinned alloc 512MB", f"{(time.perf_counter()-t0)*1e3:.1f} ms")
del x; t0 = time.perf_counter(); x = torch.empty(n, 8, pin_memory=True); print("pinned alloc again", f"{(time.perf_counter()-t0)*1e3:.1f} ms")
v, t = syn.canopy_mesh(2, 1_000_000)
sc = RaycastingScene(); sc.add_triangles(v, t); sc.commit()
lo, hi = v.min(0), v.max(0)
rays = torch.from_numpy(syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(40, 30), 4000, 4000), 4000, 4000)).pin_memory()
L = _lib.load()
outs = [torch.empty(n, dtype=torch.float32, pin_memory=True), torch.empty(n, dtype=torch.int32, pin_memory=True), torch.empty(n, dtype=torch.int32, pin_memory=True),
        torch.empty(n, 2, pin_memory=True), torch.empty(n, 3, pin_memory=True)]
P = lambda x: C.c_void_p(x.data_ptr())
for chunk in ("default",):
    for rep in range(3):
        t0 = time.perf_counter()
        _lib.check(L.qsmrt_cast_rays_host(sc._h, P(rays), n, *[P(x) for x in outs]))
        dt = time.perf_counter() - t0
    print("C ABI host path, preallocated pinned outputs:", f"{dt*1e3:.2f} ms", f"{n/dt/1e6:.0f} Mrays/s")
for rep in range(3):
    t0 = time.perf_counter(); a = sc.cast_rays(rays); dt = time.perf_counter() - t0
print("python cast_rays(pinned):", f"{dt*1e3:.2f} ms", f"{n/dt/1e6:.0f} Mrays/s")
