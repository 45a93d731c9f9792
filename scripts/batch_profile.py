"""batch_filter 4/2 fp32, 2^18 filters x 32 epochs, for an ncu capture."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.common import workloads as wl
N, T = 1 << 18, 32
w = wl.kf_bank_cv2d(N, steps=T, dtype=np.float32)
kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=False)
for k in "xPFHQR":
    setattr(kf, k, w[k])
zs = torch.from_numpy(w["zs"]).cuda()
for _ in range(3):
    kf.batch_filter(zs)
torch.cuda.synchronize()
