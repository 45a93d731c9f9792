import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200.kalman import KalmanFilter
from filterpy_b200.common import workloads as wl
from oracle import kf as okf
n, m, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dt = np.float64 if (len(sys.argv) < 5 or sys.argv[4] == "f64") else np.float32
w = wl.kf_bank_ca3d(N, seed=1, steps=1) if n == 9 else wl.kf_bank_cv2d(N, seed=1, steps=1)
kf = KalmanFilter(n, m, n_filters=N, dtype=dt, diagnostics=False)
for k in "xPFHQR":
    setattr(kf, k, w[k])
print("launching", n, m, N, flush=True)
kf.predict(); kf.update(w["zs"][0])
torch.cuda.synchronize()
o = okf.kf_step_bank(w["x"], w["P"], w["zs"][0], w["F"], w["H"], w["Q"], w["R"])
print("max rel err x", np.abs(kf.x.cpu().numpy() - o["x"]).max() / np.abs(o["x"]).max(),
      "P", np.abs(kf.P.cpu().numpy() - o["P"]).max() / np.abs(o["P"]).max(), flush=True)
