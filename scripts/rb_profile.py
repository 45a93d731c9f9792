"""A few fused steps of a 9/3 fp64 bank (config C3 shape, 500 k filters) for an ncu capture."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.common import workloads as wl
N = 500000
w = wl.kf_bank_ca3d(50000, steps=1)
r = N // 50000
kf = KalmanFilter(9, 3, n_filters=N, dtype=np.float64, diagnostics=len(sys.argv) > 1 and sys.argv[1] == "diag")
for k in "xPFHQR":
    a = w[k]
    setattr(kf, k, np.tile(a, (r,) + (1,) * (a.ndim - 1)))
z = torch.from_numpy(np.tile(w["zs"][0], (r, 1))).cuda()
for _ in range(3):
    kf.predict(); kf.update(z)
torch.cuda.synchronize()
