"""torchrun --nproc-per-node G scripts/dist_resample_check.py : sharded resample over G GPUs (NCCL)
must equal the single-array oracle; also times it.  Rank 0 prints the verdict."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.common import workloads as wl
from filterpy_b200 import distributed as bd
from oracle import resample as ors

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
N, u = 1 << lg, 0.3141592653
w = wl.resample_weights(N, "heavy", seed=97)
b = bd.shard_bounds(N, world)
w_loc = torch.from_numpy(w[int(b[rank]):int(b[rank + 1])]).cuda()
method = sys.argv[2] if len(sys.argv) > 2 else "compose"
best = 1e9
for it in range(6):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if method == "plan":
        if it == 0:
            splan = bd.ShardedResamplePlan([int(b[r + 1] - b[r]) for r in range(world)])
        idx, rng_t = splan.resample(w_loc, u)
        info = splan.info
    else:
        idx, rng_t, info, keep = bd.sharded_systematic_resample(w_loc, u, sizes=[int(b[r + 1] - b[r]) for r in range(world)], method=method)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if it >= 2:
        best = min(best, float(ms.item()))
lo, hi = [int(v) for v in rng_t.cpu().numpy()]
parts = [None] * world
dist.all_gather_object(parts, (lo, hi, idx[:hi - lo].cpu().numpy(), info.cpu().numpy().tolist()))
if rank == 0:
    out = np.full(N, -1, dtype=np.int32)
    for (l, h, arr, inf) in parts:
        out[l:h] = arr
    want = ors.systematic_resample_c(w, u)
    print("world", world, "N=2^%d" % lg, "bit-exact:", bool(np.array_equal(out, want)), "ranges", [(p[0], p[1]) for p in parts],
          "info", [p[3] for p in parts], "method", method, "max-rank ms (best of 4) %.3f" % best, flush=True)
# the step after: particles[idx], re-sharded evenly (local gather + slice exchange over NVLink)
particles = np.random.default_rng(11).normal(size=(N, 4)).astype(np.float32)
p_loc = torch.from_numpy(particles[int(b[rank]):int(b[rank + 1])]).cuda()
for it in range(3):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    new_loc = bd.redistribute_after_resample(p_loc, idx, rng_t, N)
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
want_idx = ors.systematic_resample_c(w, u)
ok_loc = bool(np.array_equal(new_loc.cpu().numpy(), particles[want_idx][int(b[rank]):int(b[rank + 1])]))
oks = [None] * world
dist.all_gather_object(oks, ok_loc)
if rank == 0:
    print("redistribute particles[idx]: exact on every rank:", all(oks), "max-rank ms %.3f" % float(ms2.item()), flush=True)
dist.destroy_process_group()
