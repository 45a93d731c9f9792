"""Multi-GPU plumbing: one process per GPU (torch.distributed; NCCL over NVLink on the GPU box,
gloo in the CPU tests).

Filter banks shard embarrassingly: every rank owns a contiguous slice of the N axis and runs the
same kernels on it; there is NO data-path collective.  Particle sets shard the same
way; the only exchange is the weight-sum all-reduce before a resample (north_star), plus — for a
resample that must equal the single-array reference bit for bit — an all-gather of the shards'
composite parity maps, from which every rank derives its exact carry locally
(see ``sharded_systematic_resample``).
"""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_of", "all_reduce_sum", "exclusive_prefix", "gather_counts",
           "sharded_systematic_resample", "ShardedResamplePlan", "exchange_plan", "exchange_rows", "redistribute_after_resample"]


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


def sharded_systematic_resample(weights_local, u, group=None, capacity=None, uniforms=None, sizes=None,
                                method="compose", local_sum=None):
    """Systematic (or, with ``uniforms`` = the replicated global U[N], stratified) resampling of a
    particle set whose weights are sharded contiguously over the ranks of ``group`` — one process
    per GPU, NCCL over NVLink.  Bit-identical to ``systematic_resample`` on the concatenated array.

    Exchanges (all stream-ordered, no host synchronisation besides the optional size gather):
      1. all-gather of the approximate shard weight sums (one double per rank) — the weight-sum
         exchange of the north star (``local_sum``: pass it when the caller already has it, e.g. from
         normalising the weights);
      2. ``method="compose"`` (default): all-gather of every shard's COMPOSITE — the short list of
         parity maps / true adds that carries the exact running sum across the shard, formed from
         the approximate carry alone — after which every rank derives its exact carry locally: no
         rank waits for another rank's chain.  ``method="relay"``: the exact running sum handed from
         rank r to rank r+1 (one double, point-to-point; serial over the ranks).  A shard whose
         composite cannot be formed (a dense zone of tiny weights next to a binade boundary) sets
         bit 8 of ``info[4]``: call again with ``method="relay"``.

    Returns ``(indexes, out_range, info, keep)``: rank r owns the global output positions
    ``[out_range[0], out_range[1])`` (device int64[2]); ``indexes[:out_range[1]-out_range[0]]`` holds
    their GLOBAL particle numbers (int32); ``info`` is the device int32[8] of ``bke.h``; ``keep``
    holds the scratch tensors the stream-ordered launches still use.  ``capacity`` bounds the local
    output buffer (default ``2 * n_local + 1024``; a shard holding more weight than that needs a
    larger one — info[6])."""
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
    with torch.cuda.device(dev):                             # the launches bind to the device of the weights
        ws_bytes = int(lib.bke_resample_workspace_bytes(n_local))
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
        idx = torch.empty(cap, dtype=torch.int32, device=dev)
        info = torch.zeros(8, dtype=torch.int32, device=dev)
        out_range = torch.zeros(2, dtype=torch.int64, device=dev)
        carry_out = torch.zeros(1, dtype=torch.float64, device=dev)
        st = stream_ptr(dev)
        if local_sum is None:
            local_sum = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.check(lib.bke_weights_sum(n_local, weights_local.data_ptr(), local_sum.data_ptr(), ws_ptr, ws_bytes, st))
        sums = gather_counts(local_sum.reshape(-1)[0], group)            # approximate shard sums, every rank
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
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        comp = allc = None
        if world > 1 and method == "compose":
            cbytes = int(lib.bke_resample_composite_bytes())
            comp = torch.empty(cbytes, dtype=torch.uint8, device=dev)
            allc = torch.empty(world * cbytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.bke_resample_shard_compose(ctypes.byref(a), comp.data_ptr(), st))
            dist.all_gather_into_tensor(allc, comp, group=group)
            _lib.check(lib.bke_resample_compose_carry(rank, allc.data_ptr(), carry_in.data_ptr(), status.data_ptr(), st))
            a.carry_exact = carry_in.data_ptr()
        elif rank > 0:
            dist.recv(carry_in, src=dist.get_global_rank(group, rank - 1) if group is not None else rank - 1, group=group)
            a.carry_exact = carry_in.data_ptr()
        a.phase = 2
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), stream_ptr(dev)))   # exact chain: produces carry_out
        if world > 1 and method != "compose" and rank < world - 1:             # the next rank can start its chain now
            dist.send(carry_out, dst=dist.get_global_rank(group, rank + 1) if group is not None else rank + 1, group=group)
        a.phase = 4
        _lib.check(lib.bke_resample_shard(ctypes.byref(a), stream_ptr(dev)))   # emit the indexes
        info[4:5] += status * 256                                             # bit 8: a composite could not be formed
    keep = (ws, carry_approx, carry_in, local_sum, sums, comp, allc, status, carry_out)
    return idx, out_range, info, keep


class ShardedResamplePlan(object):
    """Pre-allocated multi-GPU resample: one process per GPU, the weights sharded contiguously
    (``sizes`` = particles per rank).  Per call: three C-ABI calls and two NCCL all-gathers, nothing
    allocated, no host synchronisation —

        stage 1 (pass A, shard sum)  ->  all-gather of the shard sums (world doubles)
        stage 2 (passes B, C, the shard's composite)  ->  all-gather of the composites
        stage 3 (exact carry from the composites, exact chain, emit)

    No rank waits for another rank's chain (the composites depend on the approximate carry only).
    ``resample(weights_local, u)`` returns ``(indexes, out_range)`` as
    ``sharded_systematic_resample``; ``info`` / ``status`` are device tensors (bit 8 of info[4] or
    status != 0: a composite could not be formed — fall back to ``sharded_systematic_resample(...,
    method="relay")``)."""

    def __init__(self, sizes, group=None, device=None, capacity=None, uniforms=None):
        import ctypes
        from . import _lib
        self._ctypes, self._lib_mod = ctypes, _lib
        self.lib = _lib.load()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if len(sizes) != self.world:
            raise ValueError("sizes must have one entry per rank")
        self.sizes = [int(v) for v in sizes]
        self.n_local = self.sizes[self.rank]
        self.n_global = int(sum(self.sizes))
        self.j_offset = int(sum(self.sizes[:self.rank]))
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        cap = int(capacity) if capacity is not None else 2 * self.n_local + 1024
        self.capacity = cap
        lib = self.lib
        self.ws_bytes = int(lib.bke_resample_workspace_bytes(self.n_local))
        self.ws = torch.empty(self.ws_bytes + 256, dtype=torch.uint8, device=dev)
        self.ws_ptr = self.ws.data_ptr() + ((-self.ws.data_ptr()) % 256)
        self.indexes = torch.empty(cap, dtype=torch.int32, device=dev)
        self.info = torch.zeros(8, dtype=torch.int32, device=dev)
        self.out_range = torch.zeros(2, dtype=torch.int64, device=dev)
        self.carry_out = torch.zeros(1, dtype=torch.float64, device=dev)
        self.local_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.all_sums = torch.zeros(self.world, dtype=torch.float64, device=dev)
        cbytes = int(lib.bke_resample_composite_bytes())
        self.comp = torch.zeros(cbytes, dtype=torch.uint8, device=dev)
        self.all_comp = torch.zeros(self.world * cbytes, dtype=torch.uint8, device=dev)
        self.scratch = torch.zeros(2, dtype=torch.float64, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.uniforms = uniforms
        a = _lib.ResampleShardArgs()
        a.n_local, a.n_global, a.j_offset, a.capacity = self.n_local, self.n_global, self.j_offset, cap
        a.uniforms = None if uniforms is None else uniforms.data_ptr()
        a.indexes, a.out_range, a.carry_out = self.indexes.data_ptr(), self.out_range.data_ptr(), self.carry_out.data_ptr()
        a.workspace, a.workspace_bytes, a.info = self.ws_ptr, self.ws_bytes, self.info.data_ptr()
        a.is_last = 1 if self.rank == self.world - 1 else 0
        a.phase = 7
        self._a = a
        e = _lib.ResampleShardExt()
        e.shard_sum_out = self.local_sum.data_ptr()
        e.shard_sums_all = self.all_sums.data_ptr()
        e.composite_out = self.comp.data_ptr()
        e.composites_all = self.all_comp.data_ptr()
        e.carry_approx_buf = self.scratch.data_ptr()
        e.carry_exact_buf = self.scratch.data_ptr() + 8
        e.compose_status = self.status.data_ptr()
        e.shard_rank, e.n_shards = self.rank, self.world
        self._e = e

    def resample(self, weights_local, u):
        from ._dev import stream_ptr
        if not (weights_local.is_cuda and weights_local.dtype == torch.float64 and weights_local.is_contiguous()
                and weights_local.numel() == self.n_local):
            raise ValueError("weights_local must be a contiguous float64 CUDA tensor of %d elements" % self.n_local)
        ct, lm, lib = self._ctypes, self._lib_mod, self.lib
        a, e = self._a, self._e
        a.weights = weights_local.data_ptr()
        a.u = float(u)
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            lm.check(lib.bke_resample_shard_stage(ct.byref(a), ct.byref(e), 1, st))
            if self.world > 1:
                dist.all_gather_into_tensor(self.all_sums, self.local_sum, group=self.group)
            else:
                self.all_sums.copy_(self.local_sum)
            lm.check(lib.bke_resample_shard_stage(ct.byref(a), ct.byref(e), 2, st))
            if self.world > 1:
                dist.all_gather_into_tensor(self.all_comp, self.comp, group=self.group)
            else:
                self.all_comp.copy_(self.comp)
            lm.check(lib.bke_resample_shard_stage(ct.byref(a), ct.byref(e), 3, st))
        return self.indexes, self.out_range


# --------------------------------------------------------------------------- particle re-sharding
def exchange_plan(out_ranges, bounds, rank):
    """Who sends what after a sharded resample (the gather ``particles[indexes]``
    that follows a resample crosses shards).

    Rank r emitted the indexes of the output positions ``out_ranges[r] = [o_r, o_r+1)``; every one
    of them points into r's OWN particle shard, so r can gather those rows locally.  The new
    particle set is sharded evenly again by ``bounds`` (``shard_bounds(n, world)``), so r must ship
    the rows of positions ``[max(o_r, b_d), min(o_r+1, b_d+1))`` to rank d.  Returns
    ``(sends, recvs)``: ``sends`` = list of (dst, lo, hi) in positions relative to o_r,
    ``recvs`` = list of (src, lo, hi) relative to b_rank — both ordered by peer rank."""
    world = len(bounds) - 1
    o_lo, o_hi = int(out_ranges[rank][0]), int(out_ranges[rank][1])
    sends, recvs = [], []
    for d in range(world):
        lo, hi = max(o_lo, int(bounds[d])), min(o_hi, int(bounds[d + 1]))
        if hi > lo:
            sends.append((d, lo - o_lo, hi - o_lo))
    b_lo, b_hi = int(bounds[rank]), int(bounds[rank + 1])
    for s_ in range(world):
        lo, hi = max(int(out_ranges[s_][0]), b_lo), min(int(out_ranges[s_][1]), b_hi)
        if hi > lo:
            recvs.append((s_, lo - b_lo, hi - b_lo))
    return sends, recvs


def exchange_rows(rows, out, sends, recvs, group=None):
    """Point-to-point all-to-all-v of contiguous row slices: ``rows[lo:hi]`` to every (dst, lo, hi)
    of ``sends``, ``out[lo:hi]`` from every (src, lo, hi) of ``recvs``; slices for this rank itself
    are copied.  Works on NCCL (GPU tensors, over NVLink) and gloo (CPU tensors, tests)."""
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0

    def peer(r):
        return dist.get_global_rank(group, r) if group is not None else r
    ops = []
    self_send = [x for x in sends if x[0] == rank]
    self_recv = [x for x in recvs if x[0] == rank]
    for (_, slo, shi), (_, rlo, rhi) in zip(self_send, self_recv):
        out[rlo:rhi].copy_(rows[slo:shi])
    for src, lo, hi in recvs:
        if src != rank:
            ops.append(dist.P2POp(dist.irecv, out[lo:hi], peer(src), group))
    for dst, lo, hi in sends:
        if dst != rank:
            ops.append(dist.P2POp(dist.isend, rows[lo:hi], peer(dst), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def redistribute_after_resample(particles_local, idx_local, out_range, n_global, group=None, j_offset=None):
    """The step after a sharded resample, on the GPU: ``new_particles = particles[indexes]`` with
    the result sharded evenly again.  ``particles_local`` is this rank's shard (rows b_r..b_r+1),
    ``idx_local`` / ``out_range`` what ``sharded_systematic_resample`` returned (global particle
    numbers of the output positions ``[out_range[0], out_range[1])``).  Local gather with
    ``bke_gather_rows``, then one point-to-point exchange of contiguous slices."""
    from .monte_carlo import gather_particles
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    bounds = shard_bounds(n_global, world)
    rng_t = out_range if isinstance(out_range, torch.Tensor) else torch.as_tensor(out_range)
    if world > 1:
        allr = [torch.zeros_like(rng_t) for _ in range(world)]
        dist.all_gather(allr, rng_t.contiguous(), group=group)
        out_ranges = [tuple(int(v) for v in t.tolist()) for t in allr]
    else:
        out_ranges = [tuple(int(v) for v in rng_t.tolist())]
    cnt = out_ranges[rank][1] - out_ranges[rank][0]
    # ``j_offset``: global number of this rank's first particle — differs from bounds[rank] when the
    # weights were sharded with custom ``sizes``; the NEW set is always sharded evenly by ``bounds``
    j0 = int(bounds[rank]) if j_offset is None else int(j_offset)
    if int(particles_local.shape[0]) != int(bounds[rank + 1] - bounds[rank]) and j_offset is None:
        raise ValueError("particles_local does not match shard_bounds(n_global): pass j_offset for custom shard sizes")
    local_idx = (idx_local[:cnt].to(torch.int64) - j0).to(torch.int32)
    rows = gather_particles(particles_local, local_idx)                 # every index is in this rank's shard
    out = torch.empty((int(bounds[rank + 1] - bounds[rank]),) + tuple(particles_local.shape[1:]),
                      dtype=particles_local.dtype, device=particles_local.device)
    sends, recvs = exchange_plan(out_ranges, bounds, rank)
    return exchange_rows(rows, out, sends, recvs, group)
