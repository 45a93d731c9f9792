"""Timing sweep of the resample kernel variants: python scripts/rs_sweep.py log2N kind combo [combo ...]
combo = impl:nw:ctas:sleep_ns[:stages]   e.g. new:8:2:0:2   (env knobs BKE_RS_IMPL / _WARPS / _CTAS / _SLEEP)"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.common import workloads as wl
from filterpy_b200.monte_carlo import ResamplePlan

lg = int(sys.argv[1]); kind = sys.argv[2]; combos = sys.argv[3:]
check = os.environ.get("RS_CHECK", "1") == "1"
N = 1 << lg
w = wl.resample_weights(N, kind, seed=97)
wd = torch.from_numpy(w).cuda()
ref = None
for c in combos:
    impl, nw, ctas, sl, st = (c.split(":") + ["2"])[:5]
    os.environ["BKE_RS_STAGES"] = st
    os.environ["BKE_RS_IMPL"] = impl; os.environ["BKE_RS_WARPS"] = nw; os.environ["BKE_RS_CTAS"] = ctas
    os.environ["BKE_RS_SLEEP"] = sl
    plan = ResamplePlan(N)
    for _ in range(3):
        plan.systematic(wd, 0.0763)
    torch.cuda.synchronize()
    reps = int(os.environ.get("RS_REPS", "10"))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        plan.systematic(wd, 0.0763)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    same = None
    if check:
        idx = plan.indexes.clone()
        if ref is None:
            ref = idx
        same = bool((idx == ref).all())
    print("SWEEP 2^%d %s %s min=%.3f med=%.3f ms GB/s(12B)=%.0f frac=%.3f same_as_first=%s info=%s" % (
        lg, kind, c, min(ms), sorted(ms)[len(ms) // 2], 12.0 * N / (min(ms) * 1e-3) / 1e9,
        12.0 * N / (min(ms) * 1e-3) / 6571.6e9, same, plan.info().tolist()), flush=True)
    del plan
