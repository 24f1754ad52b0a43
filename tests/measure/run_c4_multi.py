"""BASELINE config 4 on N GPUs: "multi-tree plot mesh ~50M triangles (LBVH build-dominated), rays sharded across
8 x B200 with NCCL mesh broadcast".  Rank 0 makes the mesh, NCCL broadcasts it (1.8 GB), every rank builds the
identical LBVH, the 16M-ray grid is dealt in blocks of 4 rows round-robin to the ranks, per-ray results are gathered
on rank 0 and compared with rank 0 casting the whole grid alone.  Measurement script (one JSON line):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/measure/run_c4_multi.py
"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.distributed as dist
from pyqsm_b200 import RaycastingScene, synthetic as syn, _lib
from pyqsm_b200.distributed import broadcast_mesh

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L = _lib.load()
P = lambda x: C.c_void_p(x.data_ptr()); F3 = lambda x: (C.c_float * 3)(*[float(y) for y in x])
G, RB = 4000, 4
n_can = int(os.environ.get("C4_CANOPIES", "25"))
if rank == 0:
    v_np, t_np = syn.plot_mesh(4, n_can, 1_000_000, 40.0)
    v, t = torch.from_numpy(v_np).to(dev), torch.from_numpy(t_np.view(np.int32)).to(dev)
else:
    v = t = None
bcast_ms = 0.0
if world > 1:
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    v, t = broadcast_mesh(v, t, src=0, device=dev)
    torch.cuda.synchronize(); dist.barrier()
    bcast_ms = (time.perf_counter() - t0) * 1e3
s = RaycastingScene(device=dev, output_device=dev)
s.add_triangles(v, t.view(torch.uint32))
build_ms = s.commit()
st = s.stats()
lo, hi = np.asarray(st["scene_lo"], np.float64), np.asarray(st["scene_hi"], np.float64)
g = syn.parallel_ray_grid(lo, hi, syn.sun_direction(60, 30), G, G)
full = torch.empty(G * G, 6, dtype=torch.float32, device=dev)
_lib.check(L.qsmrt_gen_parallel_rays(P(full), G, G, F3(g[0]), F3(g[1]), F3(g[2]), F3(g[3]), None))
nb = G // RB // world
mine = full.view(nb, world, RB * G, 6)[:, rank].contiguous().view(-1, 6)           # my row blocks: [nb * RB * G, 6]
rows = mine.shape[0] // G
out = [torch.empty(mine.shape[0], dtype=torch.float32, device=dev), torch.empty(mine.shape[0], dtype=torch.uint32, device=dev)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def cast(r, w, h, o):
    _lib.check(L.qsmrt_cast_rays_2d(s._h, P(r), w, h, P(o[0]), None, P(o[1]), None, None, None))
cast(mine, G, rows, out); torch.cuda.synchronize()
if world > 1: dist.barrier()
e0.record(); cast(mine, G, rows, out); e1.record(); torch.cuda.synchronize()
cast_ms = e0.elapsed_time(e1)
# gather per-ray results on rank 0 (NCCL has no uint32: ids travel as int32 bit patterns); first a small warm-up
# gather: NCCL sets up its point-to-point channels on the first call
if world > 1:
    w = torch.zeros(8, device=dev)
    dist.gather(w, [torch.empty_like(w) for _ in range(world)] if rank == 0 else None, dst=0)
    torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
if world > 1:
    bt = [torch.empty_like(out[0]) for _ in range(world)] if rank == 0 else None
    bp = [torch.empty(out[1].shape, dtype=torch.int32, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(out[0], bt, dst=0); dist.gather(out[1].view(torch.int32), bp, dst=0)
torch.cuda.synchronize()
gather_ms = (time.perf_counter() - t0) * 1e3
vals = torch.tensor([build_ms, cast_ms], dtype=torch.float64, device=dev)
allv = [torch.empty_like(vals) for _ in range(world)]
if world > 1: dist.all_gather(allv, vals)
else: allv = [vals]
if rank == 0:
    parity = "n/a (single GPU)"
    if world > 1:
        ref = [torch.empty(G * G, dtype=torch.float32, device=dev), torch.empty(G * G, dtype=torch.uint32, device=dev)]
        cast(full, G, G, ref)
        got_t = torch.stack(bt, 0).view(world, nb, RB * G).permute(1, 0, 2).reshape(-1)     # back to grid order
        got_p = torch.stack(bp, 0).view(world, nb, RB * G).permute(1, 0, 2).reshape(-1)
        parity = "bit-identical" if torch.equal(got_t, ref[0]) and torch.equal(got_p, ref[1].view(torch.int32)) else "MISMATCH"
    b = torch.stack(allv).cpu().numpy()
    print(json.dumps({"config": "C4 plot %d canopies" % n_can, "n_gpus": world, "triangles": int(st["num_triangles"]), "rays": G * G,
                      "mesh_broadcast_ms": bcast_ms, "mesh_bytes": int(v.numel() * 4 + t.numel() * 4),
                      "broadcast_gbs": (v.numel() * 4 + t.numel() * 4) / max(bcast_ms, 1e-9) / 1e6 if world > 1 else None,
                      "build_ms_per_rank": [float(x) for x in b[:, 0]], "cast_ms_per_rank": [float(x) for x in b[:, 1]],
                      "mrays_s": G * G / float(b[:, 1].max()) / 1e3, "gather_ms": gather_ms, "bvh_height": int(st["bvh_height"]),
                      "multi_gpu_parity": parity}), flush=True)
if world > 1:
    dist.destroy_process_group()
