"""Config C2 with diagnostics=True (all optional outputs written), a few steps, for an ncu capture."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.common import workloads as wl
N = 1 << 20
w = wl.kf_bank_cv2d(N, steps=1, dtype=np.float32)
kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=True)
for k in "xPFHQR":
    setattr(kf, k, w[k])
z = torch.from_numpy(w["zs"][0]).cuda()
for _ in range(4):
    kf.predict(); kf.update(z)
torch.cuda.synchronize()
