import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.common import workloads as wl
N = 1 << 20
w = wl.kf_bank_cv2d(N, seed=77, steps=1, dtype=np.float32)
for trial in range(6):
    kf = KalmanFilter(4, 2, n_filters=N, dtype=np.float32, diagnostics=False)
    for k in "xPFHQR":
        setattr(kf, k, w[k])
    P0 = kf.P.clone(); x0 = kf.x.clone()
    valid = np.zeros(N, dtype=bool)
    kf.update(w["zs"][0], valid=valid)      # update-only, nobody has a measurement: state must not change
    torch.cuda.synchronize()
    badP = (kf.P != P0).any(dim=2).any(dim=1); badx = (kf.x != x0).any(dim=1)
    idx = torch.nonzero(badP).flatten().cpu().numpy()
    print("trial", trial, "bad P filters:", idx.size, "bad x:", int(badx.sum()), "first:", idx[:12], "mod128:", (idx[:12] % 128))
    if idx.size:
        i = int(idx[0])
        print("  P0:", P0[i].flatten().cpu().numpy()); print("  P :", kf.P[i].flatten().cpu().numpy())
