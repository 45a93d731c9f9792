"""Per-tile event trace of the single-pass resample kernel: python scripts/rs_trace.py [log2N] [kind]
events: 0 claim, 1 data landed, 2 AGG1 published, 3 chain start, 4 INCL1 published, 5 AGG2 published,
6 INCL2 published, 7 ready, 8 consumers start, 9 consumers done"""
import ctypes
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200 import _lib
from filterpy_b200.common import workloads as wl
from filterpy_b200.monte_carlo import ResamplePlan

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
kind = sys.argv[2] if len(sys.argv) > 2 else "heavy"
N = 1 << lg
nw = int(os.environ.get("BKE_RS_WARPS", "8"))
T = (N + nw * 512 - 1) // (nw * 512)
lib = _lib.load()
lib.bke_debug_resample_trace.argtypes = [ctypes.c_void_p]
lib.bke_debug_resample_trace.restype = None
w = torch.from_numpy(wl.resample_weights(N, kind, seed=97)).cuda()
plan = ResamplePlan(N)
for _ in range(3):
    plan.systematic(w, 0.0763)
torch.cuda.synchronize()
tr = torch.zeros(T * 10, dtype=torch.int64, device="cuda")
lib.bke_debug_resample_trace(ctypes.c_void_p(tr.data_ptr()))
plan.systematic(w, 0.0763)
torch.cuda.synchronize()
lib.bke_debug_resample_trace(None)
a = tr.cpu().numpy().reshape(T, 10).astype(np.float64)
a[:, 1] = a[:, 0]          # "landed" is no longer recorded
t0 = a[:, 0].min()
a = (a - t0) / 1e3          # microseconds
ok = (a > -1).all(axis=1) & (a[:, 6] > 0)
print("tiles", T, "traced fast tiles", int(ok.sum()), "span us", a[ok].max())
names = ["claim", "landed", "AGG1", "chainstart", "INCL1", "AGG2", "INCL2", "ready", "cons_start", "cons_done"]
d = np.diff(a[ok], axis=1)
print("mean us between consecutive events:")
for i in range(9):
    print("  %-10s -> %-10s mean %7.2f  median %7.2f  p90 %7.2f  max %7.2f" % (names[i], names[i + 1], d[:, i].mean(), np.median(d[:, i]), np.percentile(d[:, i], 90), d[:, i].max()))
# frontier speed: INCL2 time vs tile index
idx = np.flatnonzero(ok)
for ev in (2, 4, 5, 6, 8):
    tt = a[idx, ev]
    print("event %-10s: first %.1f us, last %.1f us; tiles/us overall %.1f; monotone violations %d" % (
        names[ev], tt.min(), tt.max(), len(idx) / (tt.max() - tt.min()), int((np.diff(tt) < 0).sum())))
# how far ahead of the INCL2 frontier is AGG2 published?  lag between AGG2(t) and INCL2(t-1)
lag = a[idx[1:], 5] - a[idx[:-1], 6]
print("AGG2(t) - INCL2(t-1): mean %.2f us, median %.2f, p10 %.2f, p90 %.2f (positive = the frontier had to wait for this tile's map)" % (
    lag.mean(), np.median(lag), np.percentile(lag, 10), np.percentile(lag, 90)))
lag1 = a[idx[1:], 2] - a[idx[:-1], 4]
print("AGG1(t) - INCL1(t-1): mean %.2f us, median %.2f, p10 %.2f, p90 %.2f" % (lag1.mean(), np.median(lag1), np.percentile(lag1, 10), np.percentile(lag1, 90)))
step = np.diff(a[idx, 6])
print("INCL2 frontier step per tile: mean %.3f us median %.3f p99 %.3f max %.3f; big steps (>2us): %d" % (
    step.mean(), np.median(step), np.percentile(step, 99), step.max(), int((step > 2).sum())))
big = np.argsort(-step)[:25]
print("largest INCL2 steps at tiles:", sorted(idx[big + 1].tolist()))
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace.npy"), a)
