"""Multi-GPU plumbing: one process per GPU (torch.distributed; NCCL over NVLink on the GPU box,
gloo in the CPU tests).

Filter banks shard embarrassingly: every rank owns a contiguous slice of the N axis and runs the
same kernels on it; there is NO data-path collective.  Particle sets shard the same
way; the only exchange is the weight-sum all-reduce before a resample (north_star), plus — for a
resample that must equal the single-array reference bit for bit — the exact running sum handed
from shard r to shard r+1 (see ``sharded_systematic_resample``).
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_of", "all_reduce_sum", "exclusive_prefix", "gather_counts"]


def shard_bounds(n, world):
    """Contiguous split of ``n`` units over ``world`` ranks: (world+1,) int64 offsets.  The first
    ``n % world`` ranks get one extra unit."""
    base, extra = divmod(int(n), int(world))
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:extra] += 1
    out = np.zeros(world + 1, dtype=np.int64)
    np.cumsum(sizes, out=out[1:])
    return out


def shard_of(arr, rank, world, axis=0):
    """The slice of a global array this rank owns."""
    b = shard_bounds(arr.shape[axis], world)
    sl = [slice(None)] * arr.ndim
    sl[axis] = slice(int(b[rank]), int(b[rank + 1]))
    return arr[tuple(sl)]


def all_reduce_sum(t, group=None):
    """In-place SUM all-reduce of a tensor (the particle-weight-sum exchange)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_counts(value, group=None):
    """all-gather one scalar per rank -> 1-D tensor of length world (on ``value``'s device)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value.reshape(1).clone()
    world = dist.get_world_size(group)
    out = [torch.zeros_like(value.reshape(1)) for _ in range(world)]
    dist.all_gather(out, value.reshape(1).contiguous(), group=group)
    return torch.cat(out)


def exclusive_prefix(value, group=None):
    """Sum of ``value`` over the ranks before this one (left-to-right, deterministic order)."""
    allv = gather_counts(value, group)
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    acc = torch.zeros_like(value.reshape(1))
    for r in range(rank):
        acc = acc + allv[r]
    return acc.reshape(value.shape)
