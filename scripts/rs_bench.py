"""Stand-alone driver for profiling the resample passes: python scripts/rs_bench.py [log2N] [reps] [kind]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.common import workloads as wl
from filterpy_b200.monte_carlo import ResamplePlan

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kind = sys.argv[3] if len(sys.argv) > 3 else "heavy"
N = 1 << lg
w = torch.from_numpy(wl.resample_weights(N, kind, seed=97)).cuda()
plan = ResamplePlan(N)
for _ in range(2):
    plan.systematic(w, 0.0763)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    plan.systematic(w, 0.0763)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print("N=2^%d kind=%s ms=%s  info=%s  GB/s(12B)=%.0f" % (lg, kind, ["%.3f" % m for m in ms], plan.info().tolist(),
                                                       12.0 * N / (min(ms) * 1e-3) / 1e9))
