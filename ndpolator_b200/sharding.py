"""Query-sharded data parallelism: one process per GPU, the
registered table (grid, masks, cell-major copy, kd topology) replicated on every
GPU, contiguous query ranges per rank, NO collective on the compute path -- only
an optional final gather of the results.

torch.distributed is plumbing here: NCCL over NVLink on the GPU box, gloo in the
CPU tests."""
from __future__ import annotations

import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1-process default)."""
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0)))


def shard_range(n: int, rank: int, world_size: int):
    """Contiguous range [lo, hi) of rank `rank`: [g*N/G, (g+1)*N/G)."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def ndpolate_sharded(ndpolate_local, n_total: int, make_queries, gather: bool = True, group=None):
    """Runs `ndpolate_local(q_shard) -> (interps, dists)` on this rank's contiguous share
    of an n_total-query batch (`make_queries(lo, hi)` materialises rows [lo, hi) where the
    rank wants them: host array or device tensor).  With gather=True rank 0 receives the
    full (interps, dists) in query order (torch tensors on the backend's device); other
    ranks get (None, None).  Without a process group it is a plain local call."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return ndpolate_local(make_queries(0, n_total))
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(n_total, rank, ws)
    interps, dists = ndpolate_local(make_queries(lo, hi))
    if not gather:
        return interps, dists
    interps = torch.as_tensor(interps)
    dists = torch.as_tensor(dists)
    if dist.get_backend(group) == 'nccl':
        interps, dists = interps.cuda(), dists.cuda()
    vdim = interps.shape[1]
    sizes = [shard_range(n_total, r, ws) for r in range(ws)]
    most = max(h - l for l, h in sizes)          # gather wants equal shapes: pad the short shards

    def padded(t, width):
        buf = torch.zeros((most, width), dtype=t.dtype, device=t.device)
        buf[:t.shape[0]] = t.reshape(-1, width)
        return buf

    out_i = out_d = None
    if rank == 0:
        out_i = [torch.empty((most, vdim), dtype=interps.dtype, device=interps.device) for _ in sizes]
        out_d = [torch.empty((most, 1), dtype=dists.dtype, device=dists.device) for _ in sizes]
    dist.gather(padded(interps, vdim), out_i, dst=0, group=group)
    dist.gather(padded(dists, 1), out_d, dst=0, group=group)
    if rank == 0:
        return (torch.cat([t[:h - l] for t, (l, h) in zip(out_i, sizes)]),
                torch.cat([t[:h - l] for t, (l, h) in zip(out_d, sizes)]))
    return None, None
