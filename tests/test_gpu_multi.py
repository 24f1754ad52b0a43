"""Two-GPU test of the multi-GPU path over NCCL (skipped on a single-GPU box; the same plumbing is
covered on CPU with gloo in test_distributed.py): mesh broadcast, per-rank build, sharded rays /
angles / sky directions, gather and all-reduce -- results equal the single-GPU run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from pyqsm_b200 import RaycastingScene, synthetic as syn, environment as env
    from pyqsm_b200.distributed import broadcast_mesh, cast_rays_sharded

    if rank == 0:
        v_np, t_np = syn.qsm_tree_mesh(seed=5, n_cylinders=60)
        v, t = torch.from_numpy(v_np).to(dev), torch.from_numpy(t_np.view(np.int32)).to(dev)
    else:
        v = t = None
    v, t = broadcast_mesh(v, t, src=0, device=dev)
    scene = RaycastingScene(device=dev, output_device=dev)
    scene.add_triangles(v, t.view(torch.uint32))
    st = scene.stats()
    scene.commit()
    st = scene.stats()
    lo, hi = np.asarray(st["scene_lo"]), np.asarray(st["scene_hi"])
    rays = torch.from_numpy(syn.materialize_grid(*syn.parallel_ray_grid(lo, hi, syn.sun_direction(40, 70), 301, 207), 301, 207)).to(dev)
    full, _ = cast_rays_sharded(scene.cast_rays, rays)
    angles = [(20.0, 10.0), (45.0, 100.0), (70.0, 250.0), (30.0, 300.0), (60.0, 45.0)]
    expo = env.sun_exposure(scene, angles, grid=(300, 200), shard=(rank, world))
    pts = v[::7].contiguous()
    gap = env.sky_gap_fraction(scene, pts, None, n_dirs=33, seed=3, offset=0.0, shard=(rank, world))
    if rank == 0:
        single = scene.cast_rays(rays)
        ok = all(torch.equal(full[k], single[k]) for k in single)
        expo1 = env.sun_exposure(scene, angles, grid=(300, 200))
        ok = ok and torch.equal(expo["counts"], expo1["counts"]) and int(expo1["counts"].sum()) > 1000
        ok = ok and torch.allclose(expo["cell_area"], expo1["cell_area"])
        gap1 = env.sky_gap_fraction(scene, pts, None, n_dirs=33, seed=3, offset=0.0)
        ok = ok and torch.equal(gap, gap1)
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_nccl(tmp_path):
    out = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, 29700 + (os.getpid() % 200), str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
