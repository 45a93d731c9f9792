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

__all__ = ["shard_bounds", "shard_of", "all_reduce_sum", "exclusive_prefix", "gather_counts",
           "sharded_systematic_resample"]


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


def sharded_systematic_resample(weights_local, u, group=None, capacity=None, uniforms=None, sizes=None):
    """Systematic (or, with ``uniforms`` = the replicated global U[N], stratified) resampling of a
    particle set whose weights are sharded contiguously over the ranks of ``group`` — one process
    per GPU, NCCL over NVLink.  Bit-identical to ``systematic_resample`` on the concatenated array.

    Exchanges (all stream-ordered, no host synchronisation besides the size gather):
      1. all-gather of the shard sizes (host ints) and of the approximate shard weight sums (one
         double per rank) — the weight-sum exchange of the north star;
      2. the exact running sum handed from rank r to rank r+1 (one double, point-to-point): the only
         serial dependency; the heavy passes (tile sums, parity maps) run before it on every rank.

    Returns ``(indexes, out_range)``: rank r owns the global output positions
    ``[out_range[0], out_range[1])`` (device int64[2]); ``indexes[:out_range[1]-out_range[0]]`` holds
    their GLOBAL particle numbers (int32).  ``capacity`` bounds the local output buffer (default
    ``2 * n_local + 1024``; a shard holding more weight than that needs a larger one — info[6])."""
    import ctypes
    from . import _lib
    from ._dev import stream_ptr
    lib = _lib.load()
    dev = weights_local.device
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_local = int(weights_local.numel())
    if sizes is None:                                        # pass the shard sizes when they are known: saves a host gather
        sizes = [None] * world
        if world > 1:
            dist.all_gather_object(sizes, n_local, group=group)
        else:
            sizes = [n_local]
    n_global = int(sum(sizes)); j_offset = int(sum(sizes[:rank]))
    cap = int(capacity) if capacity is not None else 2 * n_local + 1024
    ws_bytes = int(lib.bke_resample_workspace_bytes(n_local))
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    idx = torch.empty(cap, dtype=torch.int32, device=dev)
    info = torch.zeros(8, dtype=torch.int32, device=dev)
    out_range = torch.zeros(2, dtype=torch.int64, device=dev)
    carry_out = torch.zeros(1, dtype=torch.float64, device=dev)
    local_sum = torch.zeros(1, dtype=torch.float64, device=dev)
    st = stream_ptr(dev)
    _lib.check(lib.bke_weights_sum(n_local, weights_local.data_ptr(), local_sum.data_ptr(), ws_ptr, ws_bytes, st))
    sums = gather_counts(local_sum[0], group)                        # approximate shard sums, every rank
    carry_approx = sums[:rank].sum().reshape(1) if rank > 0 else torch.zeros(1, dtype=torch.float64, device=dev)
    a = _lib.ResampleShardArgs()
    a.n_local, a.n_global, a.j_offset, a.capacity = n_local, n_global, j_offset, cap
    a.weights = weights_local.data_ptr()
    a.uniforms = None if uniforms is None else uniforms.data_ptr()
    a.u = float(u)
    a.carry_approx = carry_approx.data_ptr()
    a.indexes, a.out_range, a.carry_out = idx.data_ptr(), out_range.data_ptr(), carry_out.data_ptr()
    a.workspace, a.workspace_bytes, a.info = ws_ptr, ws_bytes, info.data_ptr()
    a.is_last = 1 if rank == world - 1 else 0
    a.phase = 1
    _lib.check(lib.bke_resample_shard(ctypes.byref(a), st))          # passes A-C: no dependency on other ranks
    carry_in = torch.zeros(1, dtype=torch.float64, device=dev)
    if rank > 0:
        dist.recv(carry_in, src=dist.get_global_rank(group, rank - 1) if group is not None else rank - 1, group=group)
        a.carry_exact = carry_in.data_ptr()
    a.phase = 2
    _lib.check(lib.bke_resample_shard(ctypes.byref(a), stream_ptr(dev)))   # exact chain: produces carry_out
    if rank < world - 1:                                                   # the next rank can start its chain now
        dist.send(carry_out, dst=dist.get_global_rank(group, rank + 1) if group is not None else rank + 1, group=group)
    a.phase = 4
    _lib.check(lib.bke_resample_shard(ctypes.byref(a), stream_ptr(dev)))   # emit the indexes
    keep = (ws, carry_approx, carry_in, local_sum, sums)
    return idx, out_range, info, keep
