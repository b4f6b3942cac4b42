"""Multi-GPU plumbing of the dense extraction: x-slab partition of the lattice, all-gather
of the logit slabs and min/max exchange for the iso-level (torch.distributed; NCCL over
NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md §8e.  The reference is
single-GPU (train.py:29) — there is nothing to mirror here."""
import torch
import torch.distributed as dist


def rank_world(group=None):
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    if group is False:
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def slab(nx, rank, world):
    """rows [x0, x1) of the lattice owned by `rank`: equal slabs of ceil(nx/world) rows
    (the last ranks may own fewer / none), so every slab is one contiguous, equally
    strided block of the flat (x-slowest) logit array."""
    per = (nx + world - 1) // world
    x0 = min(nx, rank * per)
    return x0, min(nx, x0 + per)


def all_gather_slabs(grid, nx, group=None):
    """in-place all-gather of the x-slabs of `grid` (nx,nx,nx)."""
    rank, world = rank_world(group)
    if world == 1:
        return grid
    per = (nx + world - 1) // world
    if per * world == nx:
        x0, x1 = slab(nx, rank, world)
        dist.all_gather_into_tensor(grid.view(-1), grid[x0:x1].reshape(-1), group=group)
    else:  # ragged tail: gather into a padded buffer
        plane = nx * nx
        buf = grid.new_empty(per * world * plane)
        x0, x1 = slab(nx, rank, world)
        mine = grid.new_zeros(per * plane)
        mine[:(x1 - x0) * plane] = grid[x0:x1].reshape(-1)
        dist.all_gather_into_tensor(buf, mine, group=group)
        grid.view(-1).copy_(buf[:nx * plane])
    return grid


def all_reduce_minmax(keys, group=None):
    """keys int32[2] = ordered-int (min, max): min-reduce [0], max-reduce [1] in one call
    by negating the first entry."""
    rank, world = rank_world(group)
    if world == 1:
        return keys
    k = keys.to(torch.int64)
    k[0] = -k[0]
    dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
    k[0] = -k[0]
    keys.copy_(k.to(torch.int32))
    return keys
