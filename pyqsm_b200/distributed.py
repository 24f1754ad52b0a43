"""Multi-GPU plumbing for the ray-casting path: replicate the scene, shard the
rays.  One process per GPU over ``torch.distributed``
(NCCL on the B200 box, gloo in the CPU tests).  The reference has no
distributed layer at all (its only parallelism is ``joblib`` over files,
``pyQSM/pipeline.py:116``); rays are independent and the scene is read-only
after commit, so the only collectives are

  * one broadcast of the mesh (vertices + indices) per scene, after which
    every rank runs the deterministic LBVH build locally, and
  * one gather (per-ray results) or all-reduce (per-triangle / per-vertex
    aggregates) per batch.

Nothing is exchanged during traversal.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of ``n`` units for ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_angles(angles, rank: int, world: int):
    """Round-robin share of a solar sweep (interleaved so hit-rate differences
    between low and high elevations spread over the ranks)."""
    return list(angles[rank::world])


def broadcast_mesh(vertices, triangles, src: int = 0, device=None):
    """Broadcast ``vertices`` f32 [V,3] and ``triangles`` int32 [T,3] (uint32 bit
    pattern; NCCL has no uint32) from ``src`` to every rank.  Non-source ranks
    pass ``None``.  Returns the two tensors on ``device``."""
    rank = dist.get_rank()
    device = torch.device(device) if device is not None else (vertices.device if vertices is not None else torch.device("cpu"))
    shape = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        shape[0], shape[1] = vertices.shape[0], triangles.shape[0]
    dist.broadcast(shape, src)
    nv, nt = int(shape[0]), int(shape[1])
    if rank != src:
        vertices = torch.empty(nv, 3, dtype=torch.float32, device=device)
        triangles = torch.empty(nt, 3, dtype=torch.int32, device=device)
    else:
        vertices = vertices.to(device=device, dtype=torch.float32).contiguous()
        triangles = triangles.to(device=device).contiguous()
        if triangles.dtype != torch.int32:
            triangles = triangles.view(torch.int32) if triangles.dtype == torch.uint32 else triangles.to(torch.int32)
    dist.broadcast(vertices, src)
    dist.broadcast(triangles, src)
    return vertices, triangles


def cast_rays_sharded(trace, rays: torch.Tensor, gather: bool = True):
    """Shard ``rays`` [N,6] contiguously over the ranks, run ``trace(shard) ->
    dict of tensors`` (``RaycastingScene.cast_rays``) on each, and all-gather
    the per-ray results back into full-length tensors on every rank.  Shards
    are padded to equal length for the collective and trimmed afterwards."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n = rays.shape[0]
    b, e = shard_range(n, rank, world)
    part = trace(rays[b:e])
    if not gather:
        return part, (b, e)
    longest = max(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world))
    out = {}
    for k, x in part.items():
        as_int = x.dtype == torch.uint32
        y = x.view(torch.int32) if as_int else x
        pad = torch.zeros((longest,) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
        pad[: y.shape[0]] = y
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        full = torch.cat([bufs[r][: shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r in range(world)])
        out[k] = full.view(torch.uint32) if as_int else full
    return out, (b, e)


def allreduce_sum(x: torch.Tensor) -> torch.Tensor:
    """Sum a per-triangle / per-vertex aggregate over the ranks (in place)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(x)
    return x
