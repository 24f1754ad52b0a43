"""World-size-2 gloo test (CPU) of the multi-GPU plumbing: mesh broadcast, ray
sharding and the final gather / all-reduce.  The per-shard trace is the CPU
oracle here (tests may use it as the stand-in checker); on the GPU box the
same functions carry RaycastingScene.cast_rays."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from pyqsm_b200 import synthetic as syn
    from pyqsm_b200.distributed import broadcast_mesh, cast_rays_sharded, allreduce_sum, shard_range, shard_angles

    if rank == 0:
        v_np, t_np = syn.qsm_tree_mesh(seed=5, n_cylinders=12)
        v, t = torch.from_numpy(v_np), torch.from_numpy(t_np.view(np.int32))
    else:
        v = t = None
    v, t = broadcast_mesh(v, t, src=0, device="cpu")
    sc = oracle.OracleScene()
    sc.add_triangles(v.numpy(), t.numpy().view(np.uint32))
    lo, hi = v.numpy().min(0), v.numpy().max(0)
    rays = torch.from_numpy(syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(40, 70), 37, 29), 37, 29))

    def trace(shard):
        a = sc.cast_rays(shard.numpy(), 1)
        return {k: torch.from_numpy(x) for k, x in a.items()}

    full, (b, e) = cast_rays_sharded(trace, rays)
    assert (b, e) == shard_range(rays.shape[0], rank, world)
    # per-triangle exposure: local accumulate + all-reduce == accumulate over all rays
    expo = torch.zeros(t.shape[0], dtype=torch.int32)
    part = trace(rays[b:e])
    pid = part["primitive_ids"].view(torch.int32).to(torch.int64)
    hit = pid >= 0
    expo.index_add_(0, pid[hit], torch.ones(int(hit.sum()), dtype=torch.int32))
    allreduce_sum(expo)
    assert shard_angles(list(range(8)), rank, world) == list(range(rank, 8, world))
    if rank == 0:
        ref = sc.cast_rays(rays.numpy(), 1)
        ok = all(np.array_equal(full[k].numpy(), ref[k]) for k in ref)
        rp = ref["primitive_ids"]
        e2 = np.bincount(rp[rp != 0xFFFFFFFF].astype(np.int64), minlength=t.shape[0])
        ok = ok and np.array_equal(expo.numpy(), e2) and int(np.isfinite(ref["t_hit"]).sum()) > 10
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_cast(tmp_path, oracle_mod):
    out = tmp_path / "result.txt"
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_shard_range_covers():
    from pyqsm_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
